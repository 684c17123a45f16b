"""lorads_b200 -- ctypes binding of the B200-native LoRADS C ABI (include/lorads_b200.h).

The product is the C-ABI shared library `lib/liblorads_b200.so` (hand-written sm_100a CUDA) and the C host
driver `lib/lorads_b200` / `lib/liblorads_host.so` built from `csrc/`.  This module only marshals numpy
arrays into those entry points; it contains no arithmetic of its own and NO fallback: importing it without
the built library, or creating a Context without a CUDA device, raises.

Method names follow the reference's vocabulary (lorads/src/src_semi): cones, constraints, factors R/U/V,
constrVal, ALMCalGrad, LBFGSDirection, ALMCalq12p12, admmUpdateVar ...
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "liblorads_b200.so")
HOST_LIB_PATH = os.path.join(LIB_DIR, "liblorads_host.so")
BINARY_PATH = os.path.join(LIB_DIR, "lorads_b200")
CSRC_DIR = os.path.join(_HERE, "csrc")

R, U, V, GRAD = 0, 1, 2, 3
VEC_DUAL, VEC_CONSTR_SUM, VEC_ARD, VEC_ADD, VEC_M1, VEC_B = 0, 1, 2, 3, 4, 5
PAIR_RR, PAIR_RU, PAIR_UU, PAIR_UV = 0, 1, 2, 3

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_lp = ctypes.POINTER(ctypes.c_int64)
_c_ip = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p


class LoradsError(RuntimeError):
    pass


def build(verbose: bool = False) -> None:
    """Compile csrc/ in-tree for sm_100a (make; nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    out = subprocess.run(["make", "-C", CSRC_DIR, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
        print(out.stderr)
    if out.returncode != 0:
        raise LoradsError("building liblorads_b200.so failed")


_lib = None
_host = None


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_dp)


def _i64(a: np.ndarray):
    return a.ctypes.data_as(_c_lp)


def lib() -> ctypes.CDLL:
    """The C-ABI library.  Raises if it has not been built: there is no Python/CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LoradsError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() or `make -C {CSRC_DIR}`; "
                          "lorads_b200 has no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    i, i64, d = ctypes.c_int, ctypes.c_int64, ctypes.c_double
    sig = {
        "lgpu_create": (i, [ctypes.POINTER(_vp), i]),
        "lgpu_destroy": (None, [_vp]),
        "lgpu_last_error": (ctypes.c_char_p, [_vp]),
        "lgpu_version": (ctypes.c_char_p, []),
        "lgpu_launch_count": (i64, [_vp]),
        "lgpu_sync": (i, [_vp]),
        "lgpu_timer_record": (i, [_vp, i]),
        "lgpu_timer_elapsed_ms": (i, [_vp, i, i, _c_dp]),
        "lgpu_profile_enable": (i, [_vp, i]),
        "lgpu_profile_read": (i, [_vp, i, _c_dp, _c_lp]),
        "lgpu_profile_num_classes": (i, []),
        "lgpu_profile_class_name": (ctypes.c_char_p, [i]),
        "lgpu_partition_rows": (i, [i64, i, i, _c_lp, _c_lp, _c_lp]),
        "lgpu_nccl_unique_id": (i, [ctypes.c_char_p]),
        "lgpu_comm_init": (i, [_vp, ctypes.c_char_p, i, i]),
        "lgpu_set_problem": (i, [_vp, i64, _c_dp, i, _c_lp, i64]),
        "lgpu_cone_upload": (i, [_vp, i, _c_lp, _c_lp, _c_dp]),
        "lgpu_lp_upload": (i, [_vp, _c_lp, _c_lp, _c_dp]),
        "lgpu_cone_info": (i, [_vp, i, _c_lp]),
        "lgpu_cone_classify": (i, [i64, i64, _c_lp, _c_lp, _c_lp]),
        "lgpu_cone_layout_build": (i, [ctypes.POINTER(_vp), i64, i64, _c_lp, _c_lp, _c_dp, i, i]),
        "lgpu_cone_layout_get": (i, [_vp, ctypes.c_char_p, i64, _vp, _c_lp, ctypes.POINTER(ctypes.c_int)]),
        "lgpu_cone_layout_free": (None, [_vp]),
        "lgpu_cone_pattern": (i, [_vp, i, i64, _c_ip, _c_ip]),
        "lgpu_constants": (i, [_vp, _c_dp]),
        "lgpu_obj_scale": (i, [_vp, d]),
        "lgpu_alloc_vars": (i, [_vp, _c_lp, i]),
        "lgpu_set_factor": (i, [_vp, i, i, _c_dp]),
        "lgpu_get_factor": (i, [_vp, i, i, _c_dp]),
        "lgpu_set_lp": (i, [_vp, i, _c_dp]),
        "lgpu_get_lp": (i, [_vp, i, _c_dp]),
        "lgpu_set_vec": (i, [_vp, i, _c_dp]),
        "lgpu_get_vec": (i, [_vp, i, _c_dp]),
        "lgpu_get_rank": (i, [_vp, i, _c_lp]),
        "lgpu_fill_factor_random": (i, [_vp, i, ctypes.c_uint64]),
        "lgpu_aug_rank": (i, [_vp, _c_lp]),
        "lgpu_init_constr_val": (i, [_vp, i]),
        "lgpu_alm_cal_grad": (i, [_vp, d, _c_dp]),
        "lgpu_lbfgs_direction": (i, [_vp, i64]),
        "lgpu_alm_linesearch_terms": (i, [_vp, d, _c_dp]),
        "lgpu_alm_step": (i, [_vp, d]),
        "lgpu_alm_inner_update": (i, [_vp, d, d, _c_dp, _c_dp]),
        "lgpu_set_fused_path": (i, [_vp, i]),
        "lgpu_uses_fused_path": (i, [_vp]),
        "lgpu_uses_peer_exchange": (i, [_vp]),
        "lgpu_cone_reorder_info": (i, [_vp, i, _c_dp]),
        "lgpu_cone_owner_map": (i, [i, _c_dp, i, ctypes.POINTER(ctypes.c_int)]),
        "lgpu_set_carried_dots": (i, [_vp, i]),
        "lgpu_set_dense_tensor_path": (i, [_vp, i]),
        "lgpu_lbfgs_push": (i, [_vp, d]),
        "lgpu_primal_infeasibility": (i, [_vp, i, _c_dp]),
        "lgpu_update_dual_var": (i, [_vp, d]),
        "lgpu_cal_obj": (i, [_vp, i, _c_dp]),
        "lgpu_cal_dual_obj": (i, [_vp, _c_dp]),
        "lgpu_alm_to_admm": (i, [_vp]),
        "lgpu_average_uv": (i, [_vp]),
        "lgpu_copy_r_to_v": (i, [_vp]),
        "lgpu_admm_update_var": (i, [_vp, d, d, i64, _c_lp]),
        "lgpu_gram": (i, [_vp, i, i, _c_dp]),
        "lgpu_dual_infeasibility": (i, [_vp, _c_dp]),
        "lgpu_op_auv": (i, [_vp, i, i64, _c_dp, _c_dp, _c_dp, _c_dp]),
        "lgpu_op_uvt": (i, [_vp, i, i64, _c_dp, _c_dp, _c_dp]),
        "lgpu_op_wsum": (i, [_vp, i, _c_dp, i, _c_dp]),
        "lgpu_op_wsum_mulrk": (i, [_vp, i, i64, _c_dp, i, _c_dp, _c_dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class _Sdpa(ctypes.Structure):
    """lh_sdpa (csrc/host/lorads_host.h)"""
    _fields_ = [("m", ctypes.c_int64), ("nBlks", ctypes.c_int64), ("blkDims", _c_lp), ("nLpCols", ctypes.c_int64),
                ("b", _c_dp), ("matBeg", ctypes.POINTER(_c_lp)), ("matIdx", ctypes.POINTER(_c_lp)),
                ("matElem", ctypes.POINTER(_c_dp)), ("lpBeg", _c_lp), ("lpIdx", _c_lp), ("lpElem", _c_dp),
                ("nElems", ctypes.c_int64)]


def host_lib() -> ctypes.CDLL:
    """The C host driver as a shared library (SDPA reader, scalar line search, whole-program entry)."""
    global _host
    if _host is not None:
        return _host
    lib()  # liblorads_host.so depends on liblorads_b200.so
    if not os.path.exists(HOST_LIB_PATH):
        raise LoradsError(f"{HOST_LIB_PATH} is missing: build it with `make -C {CSRC_DIR}`")
    H = ctypes.CDLL(HOST_LIB_PATH)
    H.lh_read_sdpa.restype = ctypes.c_int
    H.lh_read_sdpa.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Sdpa), ctypes.c_int]
    H.lh_write_sdpa_binary.restype = ctypes.c_int
    H.lh_write_sdpa_binary.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Sdpa)]
    H.lh_free_sdpa.restype = None
    H.lh_free_sdpa.argtypes = [ctypes.POINTER(_Sdpa)]
    H.lh_cubic_equation.restype = ctypes.c_int
    H.lh_cubic_equation.argtypes = [ctypes.c_double] * 4 + [_c_dp]
    H.lh_line_search.restype = ctypes.c_int
    H.lh_line_search.argtypes = [ctypes.c_double, _c_dp, _c_dp]
    H.lh_sym_eigvals.restype = ctypes.c_int
    H.lh_sym_eigvals.argtypes = [ctypes.c_int, _c_dp, _c_dp]
    H.lorads_b200_main.restype = ctypes.c_int
    H.lorads_b200_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p)]
    _host = H
    return H


class SdpaProblem:
    """An SDPA problem in the layout the reference reader produces (LReadSDPA, io/lorads_file_io.c:59):
    per SDP block a CSC over packed lower-triangular indices with m+2 column pointers (column 0 = objective,
    already negated), an optional trailing LP block as CSC over LP column ids, and b."""

    def __init__(self, m: int, dims: Sequence[int], b: np.ndarray, mat_beg: List[np.ndarray],
                 mat_idx: List[np.ndarray], mat_elem: List[np.ndarray], nlp: int = 0,
                 lp_beg: Optional[np.ndarray] = None, lp_idx: Optional[np.ndarray] = None,
                 lp_elem: Optional[np.ndarray] = None):
        self.m = int(m)
        self.dims = np.ascontiguousarray(dims, dtype=np.int64)
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.mat_beg = [np.ascontiguousarray(a, dtype=np.int64) for a in mat_beg]
        self.mat_idx = [np.ascontiguousarray(a, dtype=np.int64) for a in mat_idx]
        self.mat_elem = [np.ascontiguousarray(a, dtype=np.float64) for a in mat_elem]
        self.nlp = int(nlp)
        self.lp_beg = None if lp_beg is None else np.ascontiguousarray(lp_beg, dtype=np.int64)
        self.lp_idx = None if lp_idx is None else np.ascontiguousarray(lp_idx, dtype=np.int64)
        self.lp_elem = None if lp_elem is None else np.ascontiguousarray(lp_elem, dtype=np.float64)

    @property
    def ncones(self) -> int:
        return len(self.dims)


def cone_classify(p: "SdpaProblem", c: int) -> dict:
    """the reference's storage rules for cone c, computed on the host without a GPU (lgpu_cone_classify)"""
    o = np.zeros(6, np.int64)
    if lib().lgpu_cone_classify(int(p.dims[c]), p.m, _i64(p.mat_beg[c]), _i64(p.mat_idx[c]), _i64(o)) != 0:
        raise LoradsError("lgpu_cone_classify failed")
    return dict(nnz_rows=int(o[0]), dense_aggregate=bool(o[1]), sparse_container=bool(o[2]), nnzP=int(o[3]),
                diag_only=bool(o[4]), nnzA=int(o[5]))


LAYOUT_ARRAYS = ("pat_row", "pat_col", "cval", "c_slot", "c_coef", "a_ptr", "a_slot", "a_coef", "con_gid", "t_ptr", "t_loc",
                 "t_gid", "t_val", "f_ptr", "f_col", "f_slot", "d_row", "d_val", "mc_val", "rc_ptr", "rc_gid", "rc_a")
LAYOUT_ARRAYS_PARTITIONED = ("lf_ptr", "lf_col", "lmc_val", "lrc_ptr", "lrc_gid", "lrc_a", "send_idx", "halo_gid", "send_off",
                             "send_cnt", "recv_off", "recv_cnt")
LAYOUT_SCALARS = ("mA", "nnzP", "nnzA", "nnzC", "nnzF", "max_con_len", "max_slot_len", "c_nrm1", "c_nrm2sq", "c_nrminf",
                  "dense", "sparse_container", "diag_only", "use_halo", "halo_rows", "send_rows")


def cone_layout(p: "SdpaProblem", c: int, names: Sequence[str], world: int = 1, rank: int = 0) -> dict:
    """Device-layout arrays of cone c as lgpu_cone_upload builds them (lgpu_cone_layout_*): host code, no GPU."""
    out = {}
    L = lib()
    h = _vp()
    rc = L.lgpu_cone_layout_build(ctypes.byref(h), int(p.dims[c]), p.m, _i64(p.mat_beg[c]), _i64(p.mat_idx[c]),
                                  p.mat_elem[c].ctypes.data_as(_c_dp), int(world), int(rank))
    if rc != 0:
        raise LoradsError(f"lgpu_cone_layout_build failed with {rc}")
    try:
        cnt, eb = ctypes.c_int64(), ctypes.c_int()
        for nm in names:
            rc = L.lgpu_cone_layout_get(h, nm.encode(), 0, None, ctypes.byref(cnt), ctypes.byref(eb))
            if rc != 0:
                raise LoradsError(f"lgpu_cone_layout_get({nm}) failed with {rc}")
            a = np.zeros(cnt.value, dtype={4: np.int32, 8: np.float64, -8: np.int64}[eb.value])
            if cnt.value:
                L.lgpu_cone_layout_get(h, nm.encode(), a.nbytes, a.ctypes.data, ctypes.byref(cnt), ctypes.byref(eb))
            out[nm] = dict(zip(LAYOUT_SCALARS, a.tolist())) if nm == "scalars" else a
    finally:
        L.lgpu_cone_layout_free(h)
    return out


def cone_owner_map(cost: Sequence[float], world: int) -> np.ndarray:
    """owner rank of every cone in a by-cone partitioned run -- host logic, no GPU needed"""
    c = np.ascontiguousarray(cost, dtype=np.float64)
    out = np.zeros(len(c), dtype=np.int32)
    if lib().lgpu_cone_owner_map(len(c), c.ctypes.data_as(_c_dp), int(world), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) != 0:
        raise LoradsError("bad owner-map arguments")
    return out


def partition_rows(n: int, world: int, rank: int):
    """(lo, hi, rows_per_rank) of the row block rank `rank` owns -- host logic, no GPU needed"""
    lo, hi, rpr = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    if lib().lgpu_partition_rows(int(n), int(world), int(rank), ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(rpr)) != 0:
        raise LoradsError("bad partition arguments")
    return lo.value, hi.value, rpr.value


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    if lib().lgpu_nccl_unique_id(buf) != 0:
        raise LoradsError("ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
    return buf.raw


def read_sdpa(path: str) -> SdpaProblem:
    """Parse a .dat-s file (or its binary image) with the host reader (csrc/host/sdpa_reader.cpp)."""
    H = host_lib()
    s = _Sdpa()
    if H.lh_read_sdpa(os.fsencode(path), ctypes.byref(s), 1) != 0:
        raise LoradsError(f"cannot read SDPA file {path}")
    try:
        m, nb, nlp = int(s.m), int(s.nBlks), int(s.nLpCols)
        dims = np.ctypeslib.as_array(s.blkDims, shape=(max(nb, 1),))[:nb].copy()
        b = np.ctypeslib.as_array(s.b, shape=(m,)).copy()
        begs, idxs, elems = [], [], []
        for k in range(nb):
            beg = np.ctypeslib.as_array(s.matBeg[k], shape=(m + 2,)).copy()
            nnz = int(beg[-1])
            begs.append(beg)
            idxs.append(np.ctypeslib.as_array(s.matIdx[k], shape=(max(nnz, 1),))[:nnz].copy())
            elems.append(np.ctypeslib.as_array(s.matElem[k], shape=(max(nnz, 1),))[:nnz].copy())
        lp_beg = lp_idx = lp_elem = None
        if nlp > 0:
            lp_beg = np.ctypeslib.as_array(s.lpBeg, shape=(m + 2,)).copy()
            nnz = int(lp_beg[-1])
            lp_idx = np.ctypeslib.as_array(s.lpIdx, shape=(max(nnz, 1),))[:nnz].copy()
            lp_elem = np.ctypeslib.as_array(s.lpElem, shape=(max(nnz, 1),))[:nnz].copy()
        return SdpaProblem(m, dims, b, begs, idxs, elems, nlp, lp_beg, lp_idx, lp_elem)
    finally:
        H.lh_free_sdpa(ctypes.byref(s))


def write_sdpa_binary(path: str, p: SdpaProblem) -> None:
    """The binary image of a problem (lh_write_sdpa_binary): what reading the .dat-s text produces, array for array.
    The binary accepts it wherever it accepts a .dat-s file (recognised by its magic)."""
    H = host_lib()
    s = _Sdpa()
    nb = p.ncones
    s.m, s.nBlks, s.nLpCols = p.m, nb, p.nlp
    s.nElems = int(sum(int(b[-1]) for b in p.mat_beg) + (int(p.lp_beg[-1]) if p.nlp else 0))
    s.blkDims = _i64(p.dims)
    s.b = p.b.ctypes.data_as(_c_dp)
    begs = (_c_lp * max(nb, 1))(*[_i64(a) for a in p.mat_beg])
    idxs = (_c_lp * max(nb, 1))(*[_i64(a) for a in p.mat_idx])
    vals = (_c_dp * max(nb, 1))(*[a.ctypes.data_as(_c_dp) for a in p.mat_elem])
    s.matBeg, s.matIdx, s.matElem = begs, idxs, vals
    if p.nlp:
        s.lpBeg, s.lpIdx, s.lpElem = _i64(p.lp_beg), _i64(p.lp_idx), p.lp_elem.ctypes.data_as(_c_dp)
    if H.lh_write_sdpa_binary(os.fsencode(path), ctypes.byref(s)) != 0:
        raise LoradsError(f"cannot write {path}")


def _as_sdpa(p: SdpaProblem):
    """an lh_sdpa view of the numpy arrays of p (the returned tuple keeps the temporaries alive)"""
    s = _Sdpa()
    nb = p.ncones
    s.m, s.nBlks, s.nLpCols = p.m, nb, p.nlp
    s.nElems = int(sum(int(b[-1]) for b in p.mat_beg) + (int(p.lp_beg[-1]) if p.nlp else 0))
    s.blkDims = _i64(p.dims)
    s.b = p.b.ctypes.data_as(_c_dp)
    begs = (_c_lp * max(nb, 1))(*[_i64(a) for a in p.mat_beg])
    idxs = (_c_lp * max(nb, 1))(*[_i64(a) for a in p.mat_idx])
    vals = (_c_dp * max(nb, 1))(*[a.ctypes.data_as(_c_dp) for a in p.mat_elem])
    s.matBeg, s.matIdx, s.matElem = begs, idxs, vals
    return s, (begs, idxs, vals)


CONSTRAINT_STATS = ("fro_norm", "nnz", "trace", "diag_norm", "gershgorin", "rows_touched", "blocks_touched")


def constraint_stats(p: SdpaProblem):
    """Hand-off to the reference's feature extractor (dataset/processor.py:246-345) from the parsed problem: per-constraint
    statistics (m x 7, columns CONSTRAINT_STATS), the same seven numbers for the objective, and the constraint x row
    incidence pattern as a CSR (ptr, rows) over the block-diagonal row numbering.  Host code, no GPU."""
    H = host_lib()
    H.lh_constraint_stats.restype = ctypes.c_int
    H.lh_constraint_stats.argtypes = [ctypes.POINTER(_Sdpa), _c_dp, _c_dp]
    H.lh_constraint_rows.restype = ctypes.c_int
    H.lh_constraint_rows.argtypes = [ctypes.POINTER(_Sdpa), _c_lp, _c_lp, ctypes.POINTER(ctypes.c_int64)]
    s, keep = _as_sdpa(p)
    out = np.zeros((p.m, 7))
    obj = np.zeros(7)
    if H.lh_constraint_stats(ctypes.byref(s), out.ctypes.data_as(_c_dp), obj.ctypes.data_as(_c_dp)) != 0:
        raise LoradsError("lh_constraint_stats failed")
    ptr = np.zeros(p.m + 1, dtype=np.int64)
    cnt = ctypes.c_int64()
    if H.lh_constraint_rows(ctypes.byref(s), _i64(ptr), None, ctypes.byref(cnt)) != 0:
        raise LoradsError("lh_constraint_rows failed")
    rows = np.zeros(max(cnt.value, 1), dtype=np.int64)
    H.lh_constraint_rows(ctypes.byref(s), _i64(ptr), _i64(rows), ctypes.byref(cnt))
    del keep
    return out, obj, ptr, rows[:cnt.value]


def constraint_couplings(p: SdpaProblem) -> dict:
    """Second half of the hand-off (dataset/processor.py:347-366, :497-505, :580-600, :640-643): `cost_inner` = <A_i, C> and
    `rows_shared_with_cost` per constraint; the pairs i < j of constraints with a common row as a CSR (`ptr`, `col`) with
    `overlap` = |rows_i & rows_j| and `inner` = <A_i, A_j>.  Host code, no GPU."""
    H = host_lib()
    H.lh_constraint_cost_alignment.restype = ctypes.c_int
    H.lh_constraint_cost_alignment.argtypes = [ctypes.POINTER(_Sdpa), _c_dp, _c_lp]
    H.lh_constraint_pairs.restype = ctypes.c_int
    H.lh_constraint_pairs.argtypes = [ctypes.POINTER(_Sdpa), _c_lp, _c_lp, _c_lp, _c_dp, ctypes.POINTER(ctypes.c_int64)]
    s, keep = _as_sdpa(p)
    cost_inner = np.zeros(p.m)
    shared = np.zeros(p.m, dtype=np.int64)
    if H.lh_constraint_cost_alignment(ctypes.byref(s), cost_inner.ctypes.data_as(_c_dp), _i64(shared)) != 0:
        raise LoradsError("lh_constraint_cost_alignment failed")
    ptr = np.zeros(p.m + 1, dtype=np.int64)
    cnt = ctypes.c_int64()
    if H.lh_constraint_pairs(ctypes.byref(s), _i64(ptr), None, None, None, ctypes.byref(cnt)) != 0:
        raise LoradsError("lh_constraint_pairs failed")
    col = np.zeros(max(cnt.value, 1), dtype=np.int64)
    ov = np.zeros(max(cnt.value, 1), dtype=np.int64)
    inner = np.zeros(max(cnt.value, 1))
    if H.lh_constraint_pairs(ctypes.byref(s), _i64(ptr), _i64(col), _i64(ov), inner.ctypes.data_as(_c_dp), ctypes.byref(cnt)) != 0:
        raise LoradsError("lh_constraint_pairs failed")
    del keep
    k = cnt.value
    return {"cost_inner": cost_inner, "rows_shared_with_cost": shared, "ptr": ptr, "col": col[:k], "overlap": ov[:k], "inner": inner[:k]}


def run_solver(argv: Sequence[str], **kw) -> subprocess.CompletedProcess:
    """Run the drop-in binary exactly as benchmark.py runs the reference's (benchmark.py:240-262)."""
    if not os.path.exists(BINARY_PATH):
        raise LoradsError(f"{BINARY_PATH} is missing: build it with `make -C {CSRC_DIR}`")
    return subprocess.run([BINARY_PATH] + [str(a) for a in argv], capture_output=True, text=True, **kw)


class Context:
    """One GPU context of the C ABI (lgpu_ctx).  All arrays passed in/out are host numpy arrays; factors are
    (n, r) arrays whose Fortran-order memory equals the reference's column-major matElem."""

    def __init__(self, device: int = 0):
        self._L = lib()
        h = _vp()
        if self._L.lgpu_create(ctypes.byref(h), int(device)) != 0:
            raise LoradsError("lgpu_create: " + self._L.lgpu_last_error(None).decode())
        self._h = h
        self.m = 0
        self.dims: List[int] = []
        self.nlp = 0
        self.rank: List[int] = []

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """join the NCCL communicator of a row-block partitioned run (before load())"""
        assert len(unique_id) == 128
        self._ck(self._L.lgpu_comm_init(self._h, unique_id, int(rank), int(world)), "lgpu_comm_init")

    def reorder_info(self, cone: int = 0) -> dict:
        out = np.zeros(3)
        self._ck(self._L.lgpu_cone_reorder_info(self._h, int(cone), out.ctypes.data_as(_c_dp)), "lgpu_cone_reorder_info")
        return {"applied": bool(out[0]), "window_share_before": float(out[1]), "window_share_after": float(out[2])}

    def uses_peer_exchange(self) -> bool:
        return bool(self._L.lgpu_uses_peer_exchange(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.lgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise LoradsError(f"{what}: {self._L.lgpu_last_error(self._h).decode()}")

    # ---- problem ------------------------------------------------------------------------------
    def load(self, p: SdpaProblem) -> "Context":
        self.m, self.dims, self.nlp = p.m, [int(x) for x in p.dims], p.nlp
        self._ck(self._L.lgpu_set_problem(self._h, p.m, _dp(p.b), p.ncones, _i64(p.dims), p.nlp), "lgpu_set_problem")
        if p.nlp > 0:
            self._ck(self._L.lgpu_lp_upload(self._h, _i64(p.lp_beg), _i64(p.lp_idx), _dp(p.lp_elem)), "lgpu_lp_upload")
        for c in range(p.ncones):
            self._ck(self._L.lgpu_cone_upload(self._h, c, _i64(p.mat_beg[c]), _i64(p.mat_idx[c]), _dp(p.mat_elem[c])),
                     "lgpu_cone_upload")
        return self

    def cone_info(self, c: int) -> dict:
        o = np.zeros(6, np.int64)
        self._ck(self._L.lgpu_cone_info(self._h, c, _i64(o)), "lgpu_cone_info")
        return dict(nnz_rows=int(o[0]), dense_aggregate=bool(o[1]), sparse_container=bool(o[2]), nnzP=int(o[3]),
                    diag_only=bool(o[4]), nnzA=int(o[5]))

    def cone_pattern(self, c: int):
        k = self.cone_info(c)["nnzP"]
        row, col = np.zeros(k, np.int32), np.zeros(k, np.int32)
        self._ck(self._L.lgpu_cone_pattern(self._h, c, k, row.ctypes.data_as(_c_ip), col.ctypes.data_as(_c_ip)),
                 "lgpu_cone_pattern")
        return row, col

    def constants(self) -> np.ndarray:
        o = np.zeros(6)
        self._ck(self._L.lgpu_constants(self._h, _dp(o)), "lgpu_constants")
        return o

    def obj_scale(self, s: float):
        self._ck(self._L.lgpu_obj_scale(self._h, float(s)), "lgpu_obj_scale")

    @property
    def launch_count(self) -> int:
        return int(self._L.lgpu_launch_count(self._h))

    # ---- timing hooks ---------------------------------------------------------------------------
    def sync(self):
        self._ck(self._L.lgpu_sync(self._h), "lgpu_sync")

    def timer_record(self, slot: int):
        self._ck(self._L.lgpu_timer_record(self._h, slot), "lgpu_timer_record")

    def timer_elapsed_ms(self, a: int, b: int) -> float:
        o = ctypes.c_double()
        self._ck(self._L.lgpu_timer_elapsed_ms(self._h, a, b, ctypes.cast(ctypes.byref(o), _c_dp)), "lgpu_timer_elapsed_ms")
        return o.value

    def profile_enable(self, on: bool):
        self._ck(self._L.lgpu_profile_enable(self._h, 1 if on else 0), "lgpu_profile_enable")

    def profile_read(self) -> dict:
        k = self._L.lgpu_profile_num_classes()
        ms, cnt = np.zeros(k), np.zeros(k, np.int64)
        self._ck(self._L.lgpu_profile_read(self._h, k, _dp(ms), _i64(cnt)), "lgpu_profile_read")
        return {self._L.lgpu_profile_class_name(j).decode(): (float(ms[j]), int(cnt[j])) for j in range(k)}

    # ---- variables ----------------------------------------------------------------------------
    def alloc_vars(self, rank: Sequence[int], lbfgs_len: int = 2):
        r = np.ascontiguousarray(rank, dtype=np.int64)
        self._ck(self._L.lgpu_alloc_vars(self._h, _i64(r), int(lbfgs_len)), "lgpu_alloc_vars")
        self.rank = [int(x) for x in r]

    def set_factor(self, which: int, c: int, a: np.ndarray):
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == (self.dims[c], self.rank[c]), (a.shape, self.dims[c], self.rank[c])
        self._ck(self._L.lgpu_set_factor(self._h, which, c, _dp(a)), "lgpu_set_factor")

    def get_factor(self, which: int, c: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """`out`: optional Fortran-ordered (n, r) destination, e.g. a view of pinned memory"""
        a = np.zeros((self.dims[c], self.rank[c]), order="F") if out is None else out
        assert a.shape == (self.dims[c], self.rank[c]) and a.flags.f_contiguous and a.dtype == np.float64
        self._ck(self._L.lgpu_get_factor(self._h, which, c, _dp(a)), "lgpu_get_factor")
        return a

    def set_lp(self, which: int, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self._ck(self._L.lgpu_set_lp(self._h, which, _dp(v)), "lgpu_set_lp")

    def get_lp(self, which: int) -> np.ndarray:
        v = np.zeros(self.nlp)
        self._ck(self._L.lgpu_get_lp(self._h, which, _dp(v)), "lgpu_get_lp")
        return v

    def set_vec(self, which: int, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.m,)
        self._ck(self._L.lgpu_set_vec(self._h, which, _dp(v)), "lgpu_set_vec")

    def get_vec(self, which: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        v = np.zeros(self.m) if out is None else out
        assert v.shape == (self.m,) and v.flags.c_contiguous and v.dtype == np.float64
        self._ck(self._L.lgpu_get_vec(self._h, which, _dp(v)), "lgpu_get_vec")
        return v

    def fill_factor_random(self, which: int, seed: int):
        self._ck(self._L.lgpu_fill_factor_random(self._h, which, int(seed)), "lgpu_fill_factor_random")

    def aug_rank(self, new_rank: Sequence[int]):
        r = np.ascontiguousarray(new_rank, dtype=np.int64)
        self._ck(self._L.lgpu_aug_rank(self._h, _i64(r)), "lgpu_aug_rank")
        self.rank = [int(x) for x in r]

    # ---- lorads_func mirror -------------------------------------------------------------------
    def init_constr_val(self, pair: int):
        self._ck(self._L.lgpu_init_constr_val(self._h, pair), "lgpu_init_constr_val")

    def alm_cal_grad(self, rho: float) -> float:
        o = ctypes.c_double()
        self._ck(self._L.lgpu_alm_cal_grad(self._h, float(rho), ctypes.byref(o)), "lgpu_alm_cal_grad")
        return o.value

    def lbfgs_direction(self, inner_iter: int):
        self._ck(self._L.lgpu_lbfgs_direction(self._h, int(inner_iter)), "lgpu_lbfgs_direction")

    def alm_linesearch_terms(self, rho: float) -> np.ndarray:
        o = np.zeros(7)
        self._ck(self._L.lgpu_alm_linesearch_terms(self._h, float(rho), _dp(o)), "lgpu_alm_linesearch_terms")
        return o

    def alm_step(self, tau: float):
        self._ck(self._L.lgpu_alm_step(self._h, float(tau)), "lgpu_alm_step")

    def alm_inner_update(self, rho: float, tau: float):
        """everything of an ALM inner iteration after the line search; returns (sum |Grad|^2, pInf_l1)"""
        a, b = ctypes.c_double(), ctypes.c_double()
        self._ck(self._L.lgpu_alm_inner_update(self._h, float(rho), float(tau), ctypes.cast(ctypes.byref(a), _c_dp),
                                               ctypes.cast(ctypes.byref(b), _c_dp)), "lgpu_alm_inner_update")
        return a.value, b.value

    def set_fused_path(self, on: bool):
        self._ck(self._L.lgpu_set_fused_path(self._h, 1 if on else 0), "lgpu_set_fused_path")

    def set_dense_tensor_path(self, on: bool):
        self._ck(self._L.lgpu_set_dense_tensor_path(self._h, 1 if on else 0), "lgpu_set_dense_tensor_path")

    def set_carried_dots(self, on: bool):
        self._ck(self._L.lgpu_set_carried_dots(self._h, 1 if on else 0), "lgpu_set_carried_dots")

    @property
    def uses_fused_path(self) -> bool:
        return bool(self._L.lgpu_uses_fused_path(self._h))

    def lbfgs_push(self, tau: float):
        self._ck(self._L.lgpu_lbfgs_push(self._h, float(tau)), "lgpu_lbfgs_push")

    def primal_infeasibility(self, pair: int) -> float:
        o = ctypes.c_double()
        self._ck(self._L.lgpu_primal_infeasibility(self._h, pair, ctypes.byref(o)), "lgpu_primal_infeasibility")
        return o.value

    def update_dual_var(self, rho: float):
        self._ck(self._L.lgpu_update_dual_var(self._h, float(rho)), "lgpu_update_dual_var")

    def cal_obj(self, admm: bool) -> float:
        o = ctypes.c_double()
        self._ck(self._L.lgpu_cal_obj(self._h, 1 if admm else 0, ctypes.byref(o)), "lgpu_cal_obj")
        return o.value

    def cal_dual_obj(self) -> float:
        o = ctypes.c_double()
        self._ck(self._L.lgpu_cal_dual_obj(self._h, ctypes.byref(o)), "lgpu_cal_dual_obj")
        return o.value

    def alm_to_admm(self):
        self._ck(self._L.lgpu_alm_to_admm(self._h), "lgpu_alm_to_admm")

    def average_uv(self):
        self._ck(self._L.lgpu_average_uv(self._h), "lgpu_average_uv")

    def copy_r_to_v(self):
        self._ck(self._L.lgpu_copy_r_to_v(self._h), "lgpu_copy_r_to_v")

    def admm_update_var(self, rho: float, cg_tol: float, cg_max_iter: int, cg_iter_total: int = 0) -> int:
        o = ctypes.c_int64(cg_iter_total)
        self._ck(self._L.lgpu_admm_update_var(self._h, float(rho), float(cg_tol), int(cg_max_iter), ctypes.byref(o)),
                 "lgpu_admm_update_var")
        return int(o.value)

    def gram(self, phase: int, c: int) -> np.ndarray:
        r = self.rank[c]
        g = np.zeros((r, r))
        self._ck(self._L.lgpu_gram(self._h, phase, c, _dp(g)), "lgpu_gram")
        return g

    def dual_infeasibility(self) -> float:
        o = ctypes.c_double()
        self._ck(self._L.lgpu_dual_infeasibility(self._h, ctypes.byref(o)), "lgpu_dual_infeasibility")
        return o.value

    # ---- operator-level drop-ins on host buffers ------------------------------------------------
    def op_auv(self, c: int, Um: np.ndarray, Vm: np.ndarray):
        """LORADSUVt + coneAUV + objAUV: returns (constrVal as dense m-vector, <C, UV^T>)"""
        Um = np.asfortranarray(Um, dtype=np.float64)
        Vm = np.asfortranarray(Vm, dtype=np.float64)
        cv = np.zeros(self.m)
        o = ctypes.c_double()
        self._ck(self._L.lgpu_op_auv(self._h, c, Um.shape[1], _dp(Um), _dp(Vm), _dp(cv),
                                     ctypes.cast(ctypes.byref(o), _c_dp)), "lgpu_op_auv")
        return cv, o.value

    def op_uvt(self, c: int, Um: np.ndarray, Vm: np.ndarray) -> np.ndarray:
        Um = np.asfortranarray(Um, dtype=np.float64)
        Vm = np.asfortranarray(Vm, dtype=np.float64)
        out = np.zeros(self.cone_info(c)["nnzP"])
        self._ck(self._L.lgpu_op_uvt(self._h, c, Um.shape[1], _dp(Um), _dp(Vm), _dp(out)), "lgpu_op_uvt")
        return out

    def op_wsum(self, c: int, w: np.ndarray, add_obj: bool = True) -> np.ndarray:
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.zeros(self.cone_info(c)["nnzP"])
        self._ck(self._L.lgpu_op_wsum(self._h, c, _dp(w), 1 if add_obj else 0, _dp(out)), "lgpu_op_wsum")
        return out

    def op_wsum_mulrk(self, c: int, w: np.ndarray, X: np.ndarray, add_obj: bool = True) -> np.ndarray:
        w = np.ascontiguousarray(w, dtype=np.float64)
        X = np.asfortranarray(X, dtype=np.float64)
        Y = np.zeros(X.shape, order="F")
        self._ck(self._L.lgpu_op_wsum_mulrk(self._h, c, X.shape[1], _dp(w), 1 if add_obj else 0, _dp(X), _dp(Y)),
                 "lgpu_op_wsum_mulrk")
        return Y


# ---- synthetic instance construction (C3 / C5 shapes) ------------------------------
def pack_idx(n, i, j):
    """packed column-major lower-triangular index of (row i >= col j) (PACK_IDX, lorads_utils.h:167)"""
    return (2 * n - j - 1) * j // 2 + i


def maxcut_problem(n: int, ei: np.ndarray, ej: np.ndarray, w: np.ndarray) -> SdpaProblem:
    """MaxCut SDP of a weighted graph in the gen_MaxCut.jl file convention (lorads/data/gen_MaxCut.jl:225-229 and the
    bundled G11.dat-s): the file holds F0 = L/2 (off-diagonal -w_ij/2, diagonal +deg_i/2), A_k = e_k e_k^T, b = 1; the
    reader negates F0 (lorads_file_io.c:317-319), so LoRADS minimises <C, X> with C = -L/2 (= -2 x the MaxCut SDP value at
    the optimum; G1: -24166.3).  The arrays below hold C.  Edges must be unique with ei != ej."""
    ei = np.asarray(ei, dtype=np.int64)
    ej = np.asarray(ej, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    lo, hi = np.minimum(ei, ej), np.maximum(ei, ej)
    deg = np.zeros(n)
    np.add.at(deg, lo, w)
    np.add.at(deg, hi, w)
    dnz = np.nonzero(deg)[0].astype(np.int64)
    idx = np.concatenate([pack_idx(n, hi, lo), pack_idx(n, dnz, dnz)])
    val = np.concatenate([0.5 * w, -0.5 * deg[dnz]])
    o = np.argsort(idx, kind="stable")
    idx, val = idx[o], val[o]
    k = np.arange(n, dtype=np.int64)
    beg = np.concatenate([[0], len(idx) + np.arange(n + 1, dtype=np.int64)])
    mat_idx = np.concatenate([idx, pack_idx(n, k, k)])
    mat_elem = np.concatenate([val, np.ones(n)])
    return SdpaProblem(n, [n], np.ones(n), [beg], [mat_idx], [mat_elem])


def block_maxcut_problem(blocks: Sequence[tuple]) -> SdpaProblem:
    """A multi-block SDP: block k is the MaxCut SDP of graph k = (n, ei, ej, w); constraint ids run block after block
    (m = sum n_k) and each touches its own block only.  With more than one block the solver takes the general (per-cone)
    path, and on several GPUs the by-cone partition."""
    singles = [maxcut_problem(*b) for b in blocks]
    m = sum(q.m for q in singles)
    beg, idx, val, off = [], [], [], 0
    for q in singles:
        qb = q.mat_beg[0]
        nobj = int(qb[1])
        full = np.empty(m + 2, dtype=np.int64)
        full[0], full[1] = 0, nobj
        full[2:2 + off] = nobj                                  # constraints of earlier blocks: empty here
        full[1 + off:1 + off + q.m + 1] = qb[1:]                # this block's constraints
        full[1 + off + q.m + 1:] = qb[-1]                       # constraints of later blocks: empty here
        beg.append(full)
        idx.append(q.mat_idx[0])
        val.append(q.mat_elem[0])
        off += q.m
    return SdpaProblem(m, [int(q.dims[0]) for q in singles], np.ones(m), beg, idx, val)


def torus_graph(rows: int, cols: int, seed: int, pm1: bool = True):
    """rows x cols toroidal grid (G81-like when 100 x 200 with +-1 weights)."""
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    v = (r * cols + c).ravel()
    right = (r * cols + (c + 1) % cols).ravel()
    down = (((r + 1) % rows) * cols + c).ravel()
    ei = np.concatenate([v, v])
    ej = np.concatenate([right, down])
    keep = ei != ej
    ei, ej = ei[keep], ej[keep]
    key = np.minimum(ei, ej) * (rows * cols) + np.maximum(ei, ej)
    _, first = np.unique(key, return_index=True)
    ei, ej = ei[np.sort(first)], ej[np.sort(first)]
    rng = np.random.default_rng(seed)
    w = rng.choice([-1.0, 1.0], size=len(ei)) if pm1 else np.ones(len(ei))
    return ei, ej, w


def random_graph(n: int, out_degree: int, seed: int):
    """every vertex draws `out_degree` neighbours uniformly (average degree 2*out_degree), duplicates and
    self-loops removed, unit weights."""
    rng = np.random.default_rng(seed)
    ei = np.repeat(np.arange(n, dtype=np.int64), out_degree)
    ej = rng.integers(0, n, size=n * out_degree, dtype=np.int64)
    keep = ei != ej
    ei, ej = ei[keep], ej[keep]
    key = np.minimum(ei, ej) * n + np.maximum(ei, ej)
    key = np.unique(key)
    return key // n, key % n, np.ones(len(key))


def write_sdpa(path: str, p: SdpaProblem) -> None:
    """Write an SdpaProblem back to SDPA sparse text (objective sign restored), SDP blocks only.  Vectorised: the
    C5-sized instances have ~7e7 entries."""
    with open(path, "w") as f:
        f.write(f"{p.m}\n{p.ncones}\n{' '.join(str(int(d)) for d in p.dims)}\n")
        f.write(" ".join(repr(float(x)) for x in p.b) + "\n")
        for k in range(p.ncones):
            n = int(p.dims[k])
            beg = p.mat_beg[k]
            idx, val = p.mat_idx[k], p.mat_elem[k]
            if len(idx) == 0:
                continue
            con = np.repeat(np.arange(p.m + 1, dtype=np.int64), np.diff(beg))
            # invert pack_idx
            j = np.floor(((2 * n + 1) - np.sqrt((2.0 * n + 1) ** 2 - 8.0 * idx)) / 2.0).astype(np.int64)
            j = np.clip(j, 0, n - 1)
            j = np.where(j * (2 * n - j + 1) // 2 > idx, j - 1, j)
            j = np.where((j + 1) * (2 * n - j) // 2 <= idx, j + 1, j)
            i = idx - j * (2 * n - j + 1) // 2 + j
            v = np.where(con == 0, -val, val)
            chunk = 2_000_000
            for a in range(0, len(idx), chunk):
                sl = slice(a, a + chunk)
                cols = [con[sl].astype(str), np.full(len(con[sl]), str(k + 1)), (j[sl] + 1).astype(str),
                        (i[sl] + 1).astype(str), np.array([repr(float(x)) for x in v[sl]])]
                lines = cols[0]
                for cc in cols[1:]:
                    lines = np.char.add(np.char.add(lines, " "), cc)
                f.write("\n".join(lines.tolist()) + "\n")
