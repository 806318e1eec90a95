"""ctypes binding of libgcnb200.so (C ABI in include/gcnb200.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible the import of
the library / creation of a context raises, loudly.  Build with ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C geographconv_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgcnb200.so")

OK, E_INVALID, E_CUDA, E_WORKSPACE, E_UNSUPPORTED = 0, -1, -2, -3, -4
ACT = {"linear": 0, "tanh": 1, "relu": 2, "rectify": 2, "sigmoid": 3, "selu": 4, None: 0}
TAGS = ["spmm_a", "spmm_x", "spmm_xt", "gemm", "elementwise", "loss", "adam", "copy", "spmm_a_narrow", "comm", "sync"]
TAG_SPMM_A, TAG_SPMM_X, TAG_SPMM_XT, TAG_SPMM_A_NARROW = 0, 1, 2, 8


class GcnbError(RuntimeError):
    pass


class GcnbCsr(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int32), ("n_cols", C.c_int32), ("nnz", C.c_int64),
        ("rowptr", C.c_void_p), ("colidx", C.c_void_p), ("val", C.c_void_p),
        ("items", C.c_void_p), ("n_items", C.c_int32),
        ("long_rows", C.c_void_p), ("n_long", C.c_int32), ("n_slots", C.c_int32),
        ("tag", C.c_int32), ("engine", C.c_int32), ("unroll", C.c_int32),
    ]


class GcnbEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("act", C.c_int32), ("softmax", C.c_int32), ("accumulate", C.c_int32),
        ("dropout_p", C.c_float), ("seed", C.c_uint64), ("row0", C.c_int64), ("logits", C.c_void_p),
    ]


_i32, _i64, _u64, _f32, _vp, _sz = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_void_p, C.c_size_t
_ctxp = C.c_void_p

# name -> (restype, argtypes); every symbol include/gcnb200.h declares
SIGNATURES = {
    "gcnb_version": (C.c_int, []),
    "gcnb_create": (C.c_int, [C.c_int, _vp, C.POINTER(_ctxp)]),
    "gcnb_destroy": (C.c_int, [_ctxp]),
    "gcnb_last_error": (C.c_char_p, [_ctxp]),
    "gcnb_set_stream": (C.c_int, [_ctxp, _vp]),
    "gcnb_get_stream": (_vp, [_ctxp]),
    "gcnb_set_workspace": (C.c_int, [_ctxp, _vp, _sz]),
    "gcnb_set_option": (C.c_int, [_ctxp, C.c_char_p, C.c_int]),
    "gcnb_get_option": (C.c_int, [_ctxp, C.c_char_p, C.POINTER(C.c_int)]),
    "gcnb_sync": (C.c_int, [_ctxp]),
    "gcnb_sm_count": (C.c_int, [_ctxp]),
    "gcnb_launch_count": (C.c_longlong, [_ctxp]),
    "gcnb_prof_enable": (C.c_int, [_ctxp, C.c_int]),
    "gcnb_prof_reset": (C.c_int, [_ctxp]),
    "gcnb_prof_collect": (C.c_int, [_ctxp, C.POINTER(_f32), C.POINTER(C.c_longlong)]),
    "gcnb_h2d": (C.c_int, [_ctxp, _vp, _vp, _sz]),
    "gcnb_d2h": (C.c_int, [_ctxp, _vp, _vp, _sz]),
    "gcnb_memset": (C.c_int, [_ctxp, _vp, C.c_int, _sz]),
    "gcnb_copy2d_f32": (C.c_int, [_ctxp, _vp, _i32, _vp, _i32, _i32, _i32]),
    "gcnb_expand_u16_i32": (C.c_int, [_ctxp, _vp, _i64, _vp]),
    "gcnb_csr_plan": (C.c_int, [_vp, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _vp, _vp]),
    "gcnb_spmm_csr_f32": (C.c_int, [_ctxp, C.POINTER(GcnbCsr), _vp, _i32, _vp, _i32, _i32, C.POINTER(GcnbEpilogue)]),
    "gcnb_spmm_engine_for": (C.c_int, [_ctxp, C.POINTER(GcnbCsr), _i32, _i32]),
    "gcnb_spmm_workspace_bytes": (_sz, [C.POINTER(GcnbCsr), _i32]),
    "gcnb_gemm_f32": (C.c_int, [_ctxp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _i32]),
    "gcnb_gemm_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "gcnb_gemm_pair_f32": (C.c_int, [_ctxp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _i32]),
    "gcnb_highway_fwd_f32": (C.c_int, [_ctxp, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _i32,
                                       _vp, _i32, _vp, _i32, _vp, _i32]),
    "gcnb_highway_workspace_bytes": (_sz, [_i32, _i32]),
    "gcnb_highway_mix_f32": (C.c_int, [_ctxp, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32]),
    "gcnb_highway_bwd_f32": (C.c_int, [_ctxp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "gcnb_act_bwd_f32": (C.c_int, [_ctxp, _i32, _i32, _i32, _vp, _vp, _i32, _f32, _u64, _i64, _vp]),
    "gcnb_highway_bwd_bias_f32": (C.c_int, [_ctxp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "gcnb_act_bwd_bias_f32": (C.c_int, [_ctxp, _i32, _i32, _i32, _vp, _vp, _i32, _f32, _u64, _i64, _vp, _vp]),
    "gcnb_colsum_f32": (C.c_int, [_ctxp, _i32, _i32, _vp, _i32, _vp, _i32]),
    "gcnb_colsum_workspace_bytes": (_sz, [_i32, _i32]),
    "gcnb_csr_to_dense_f32": (C.c_int, [_ctxp, _vp, _vp, _vp, _i32, _i32, _vp, _i32]),
    "gcnb_gather_rows_f32": (C.c_int, [_ctxp, _vp, _i32, _vp, _i32, _i32, _vp, _i32]),
    "gcnb_scatter_rows_f32": (C.c_int, [_ctxp, _vp, _i32, _vp, _i32, _i32, _vp, _i32]),
    "gcnb_xent_metrics_f32": (C.c_int, [_ctxp, _vp, _i32, _i32, _vp, _vp, _i32, _vp]),
    "gcnb_xent_grad_f32": (C.c_int, [_ctxp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _f32, _vp, _i32]),
    "gcnb_xent_grad_dense_f32": (C.c_int, [_ctxp, _vp, _i32, _i32, _i32, _vp, _f32, _vp, _i32]),
    "gcnb_push_arm": (C.c_int, [_ctxp, _vp, _i32, _vp, _vp, _vp, _i64]),
    "gcnb_push_consumed": (C.c_int, [_ctxp]),
    "gcnb_gather_argmax_f32": (C.c_int, [_ctxp, _vp, _i32, _i32, _vp, _i32, _vp, _vp]),
    "gcnb_geo_distance_f64": (C.c_int, [_ctxp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, C.c_double, _vp, _vp]),
    "gcnb_l1l2_f32": (C.c_int, [_ctxp, _vp, _vp, _i64, _f32, _vp]),
    "gcnb_adam_f32": (C.c_int, [_ctxp, _vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32]),
    "gcnb_dropout_mask_u8": (C.c_int, [_ctxp, _i32, _i32, _f32, _u64, _i64, _vp]),
    "gcnb_adj_workspace_bytes": (_sz, [_i64, _i32]),
    "gcnb_adj_build_rows": (C.c_int, [_ctxp, _vp, _vp, _i64, _i32, _vp, _sz, _vp, C.POINTER(_i64)]),
    "gcnb_adj_fill_f32": (C.c_int, [_ctxp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "gcnb_adj_normalize_weighted_f64": (C.c_int, [_ctxp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "gcnb_peer_alloc": (C.c_int, [_ctxp, _sz, C.POINTER(_vp), _vp]),
    "gcnb_peer_free": (C.c_int, [_ctxp, _vp]),
    "gcnb_peer_open": (C.c_int, [_ctxp, _vp, C.POINTER(_vp)]),
    "gcnb_peer_close": (C.c_int, [_ctxp, _vp]),
    "gcnb_peer_setup": (C.c_int, [_ctxp, _i32, _i32, C.POINTER(_vp), _sz, _sz]),
    "gcnb_peer_barrier": (C.c_int, [_ctxp]),
    "gcnb_slice_push_f32": (C.c_int, [_ctxp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp]),
    "gcnb_spmm_csr_sliced_f32": (C.c_int, [_ctxp, C.POINTER(GcnbCsr), _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32,
                                           C.POINTER(GcnbEpilogue)]),
    "gcnb_row_softmax_f32": (C.c_int, [_ctxp, _vp, _i32, _i32, _i32, _vp]),
}

_lib = None


def load_library(path=None):
    """dlopen libgcnb200.so and attach prototypes.  Raises GcnbError if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise GcnbError(
            "libgcnb200.so not found at %s -- the CUDA extension is required (no CPU fallback). "
            "Build it: make -C geographconv_b200/csrc" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class Context:
    """Owns one gcnb_ctx; every wrapper raises GcnbError with the library's message on failure."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = _ctxp()
        rc = self.lib.gcnb_create(int(device), stream, C.byref(h))
        if rc != OK:
            raise GcnbError(
                "gcnb_create(device=%d) failed with %d: a B200 (sm_100) GPU is required; there is no CPU "
                "fallback" % (device, rc))
        self.h = h
        self.device = device
        # tuning knobs every context of the process honours (A/B measurements; defaults are the measured winners)
        if "GCNB_GEMM_BLO" in os.environ:  # tcgen05 GEMMs: weights' lo tile derived in shared memory (1) or loaded (0)
            self.set_option("gemm_blo", int(os.environ["GCNB_GEMM_BLO"]))
        if "GCNB_GEMM_V" in os.environ:  # tcgen05 GEMM kernel: 1 = one tile per SM (default), 2 = two co-resident CTAs per SM
            self.set_option("gemm_v", int(os.environ["GCNB_GEMM_V"]))
        if "GCNB_GEMM_BLO2" in os.environ:  # gemm_v 2: weights' residual tile derived in shared memory (1) or loaded (0, default)
            self.set_option("gemm_blo2", int(os.environ["GCNB_GEMM_BLO2"]))
        if "GCNB_GEMM_PREFETCH" in os.environ:  # gemm_v 2: L2 look-ahead of the activation rows, in 128-row tiles
            self.set_option("gemm_prefetch", int(os.environ["GCNB_GEMM_PREFETCH"]))

    def close(self):
        if getattr(self, "h", None):
            self.lib.gcnb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != OK:
            raise GcnbError("gcnb error %d: %s" % (rc, self.lib.gcnb_last_error(self.h).decode()))

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(self.h, *args))

    def set_option(self, name, value):
        self.call("gcnb_set_option", name.encode(), int(value))

    def get_option(self, name):
        v = C.c_int()
        self.call("gcnb_get_option", name.encode(), C.byref(v))
        return v.value

    def sync(self):
        self.call("gcnb_sync")

    def launch_count(self):
        return int(self.lib.gcnb_launch_count(self.h))

    def sm_count(self):
        return int(self.lib.gcnb_sm_count(self.h))

    def prof_enable(self, on=True):
        self.call("gcnb_prof_enable", 1 if on else 0)

    def prof_reset(self):
        self.call("gcnb_prof_reset")

    def prof_collect(self):
        ms = (_f32 * len(TAGS))()
        ops = (C.c_longlong * len(TAGS))()
        self.call("gcnb_prof_collect", ms, ops)
        return {t: (float(ms[i]), int(ops[i])) for i, t in enumerate(TAGS)}


def csr_plan(rowptr, chunk):
    """Host work decomposition of a CSR (items / long rows); see gcnb_csr_plan."""
    import numpy as np
    lib = load_library()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    n_rows = len(rowptr) - 1
    ni, nl, ns = _i32(), _i32(), _i32()
    rc = lib.gcnb_csr_plan(rowptr.ctypes.data, n_rows, chunk, C.byref(ni), C.byref(nl), C.byref(ns), None, None)
    if rc != OK:
        raise GcnbError("gcnb_csr_plan failed: %d" % rc)
    items = np.empty((max(ni.value, 1), 4), dtype=np.int32)
    long_rows = np.empty((max(nl.value, 1), 3), dtype=np.int32)
    rc = lib.gcnb_csr_plan(rowptr.ctypes.data, n_rows, chunk, C.byref(ni), C.byref(nl), C.byref(ns),
                           items.ctypes.data, long_rows.ctypes.data)
    if rc != OK:
        raise GcnbError("gcnb_csr_plan failed: %d" % rc)
    return items[:ni.value], long_rows[:nl.value], ns.value
