#!/usr/bin/env python
"""Time the feature-extractor hand-off (csrc/host/features.c through lorads_b200.constraint_stats / constraint_couplings,
reader included) next to the reference's own dataset/processor.py (SDPAParser + FeatureExtractor precomputation + edges)
on the same files.  CPU only; the reference leg runs only where /root/reference exists (the build container).

usage: python scripts/feature_handoff_time.py file.dat-s [...]      -> one JSON line per file
"""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
import numpy as np  # noqa: E402

import lorads_b200 as lb  # noqa: E402


def reference_leg(path):
    if not os.path.isdir("/root/reference/dataset"):
        return None
    if "torch_geometric" not in sys.modules:
        tg, tgd = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.data")
        tgd.Data = type("Data", (), {"__init__": lambda self, **kw: self.__dict__.update(kw)})
        tg.data = tgd
        sys.modules["torch_geometric"], sys.modules["torch_geometric.data"] = tg, tgd
        sys.path.insert(0, "/root/reference")
    from dataset.processor import FeatureExtractor, SDPAParser
    t0 = time.perf_counter()
    ps = SDPAParser(path)
    ps.parse()
    C, A, b, m, n, offs = ps.get_data()
    t1 = time.perf_counter()
    fe = FeatureExtractor(C, A, b, m, n, block_offsets=offs)
    t2 = time.perf_counter()
    ei, _ = fe.compute_edges()
    t3 = time.perf_counter()
    return {"parse_s": t1 - t0, "stats_and_cost_s": t2 - t1, "edges_s": t3 - t2, "total_s": t3 - t0, "edges": int(ei.shape[1] // 2)}


def main():
    for path in sys.argv[1:]:
        t0 = time.perf_counter()
        p = lb.read_sdpa(path)
        t1 = time.perf_counter()
        lb.constraint_stats(p)
        t2 = time.perf_counter()
        cp = lb.constraint_couplings(p)
        t3 = time.perf_counter()
        ours = {"parse_s": t1 - t0, "stats_and_rows_s": t2 - t1, "couplings_s": t3 - t2, "total_s": t3 - t0, "pairs": int(len(cp["col"]))}
        ref = reference_leg(path)
        print(json.dumps({"instance": os.path.basename(path), "m": int(p.m), "n": int(np.sum(p.dims)), "handoff": ours, "reference": ref,
                          "speedup_total": (ref["total_s"] / ours["total_s"]) if ref else None}), flush=True)


if __name__ == "__main__":
    main()
