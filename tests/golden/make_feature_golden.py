#!/usr/bin/env python
"""Golden vectors for the feature-extractor hand-off, made by the REFERENCE's own code: dataset/processor.py
(SDPAParser + FeatureExtractor, imported from /root/reference, unmodified) run on the committed fixture instances.

The reference imports torch_geometric only for its `Data` container, which this image does not have; a three-line stand-in
is put into sys.modules before the import -- no line of the extractor itself is replaced.  Runs in the build container only
(/root/reference does not travel); the output tests/golden/features.npz is committed.

    python tests/golden/make_feature_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
NAMES = ["theta_n30", "multiblock_lp", "multiblock_sdp", "control_like_12_6", "general_sparse_n60", "dense_constraint_n24",
         "maxcut_torus_8x10", "G11"]


def main():
    tg = types.ModuleType("torch_geometric")
    tgd = types.ModuleType("torch_geometric.data")

    class Data:  # container only
        def __init__(self, **kw):
            self.__dict__.update(kw)

    tgd.Data = Data
    tg.data = tgd
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.data"] = tgd
    sys.path.insert(0, REF)
    from dataset.processor import FeatureExtractor, SDPAParser, extract_features_only

    out = {}
    for nm in NAMES:
        path = os.path.join(HERE, "instances", nm + ".dat-s")
        ps = SDPAParser(path)
        ps.parse()
        C, A, b, m, n, offs = ps.get_data()
        fe = FeatureExtractor(C, A, b, m, n, block_offsets=offs)
        ei, ea = fe.compute_edges()
        g, x, ea2 = extract_features_only(path)
        assert np.array_equal(ea, ea2)
        out[nm + "/global"] = g
        out[nm + "/node"] = x
        out[nm + "/edge_index"] = ei
        out[nm + "/edge_attr"] = ea
        # the float64 internals the float32 features are made of
        for key in ("norms", "nnz_counts", "traces", "diag_norms", "gershgorin_bounds", "blocks_touched", "row_sizes", "cos_with_C"):
            out[nm + "/" + key] = np.asarray(getattr(fe, key))
        out[nm + "/C_frob"] = np.float64(fe.C_frob)
        out[nm + "/C_rows"] = np.asarray(fe.C_row_indices, dtype=np.int64)
        out[nm + "/n"] = np.int64(n)
        print(f"{nm:22s} m={m:4d} n={n:4d} blocks={len(offs) - 1} edges={ei.shape[1] // 2}")
    np.savez_compressed(os.path.join(HERE, "features.npz"), **out)


if __name__ == "__main__":
    main()
