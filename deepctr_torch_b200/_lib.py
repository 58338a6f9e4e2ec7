"""ctypes binding of libctr_b200.so (the C ABI declared in include/ctr_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
The product never computes the hot path on the CPU or through stock torch ops.
"""
from __future__ import annotations

import ctypes
import os
import threading

from ._build import LIB_PATH

c_i32, c_i64, c_f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
c_int, c_ptr = ctypes.c_int, ctypes.c_void_p

ABI_VERSION = 2

# name -> argtypes (restype is int for everything except the three listed below)
_P = c_ptr
SIGNATURES = {
    "ctr_gather_fwd": [_P, c_i64, c_i64, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P, c_int, _P,
                       c_int, _P, _P, _P, c_i64, _P, _P, _P, c_int, c_int, _P],
    "ctr_fm_fwd": [_P, c_i64, c_i64, c_int, c_int, _P, _P],
    "ctr_fm_bwd": [_P, c_i64, c_i64, c_int, c_int, _P, _P, c_i64, _P],
    "ctr_scatter_bwd_dense": [_P, c_i64, c_i64, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P,
                              _P, c_i64, _P, c_i64, _P, _P, c_int, _P],
    "ctr_write_ptrs": [_P, c_int, _P, _P],
    "ctr_unique_plan": [_P, c_i64, c_i64, c_int, _P, _P, _P, _P, c_i64, _P, _P, _P, _P, _P, c_int, _P, _P],
    "ctr_scatter_bwd_rowwise": [c_i64, c_int, _P, _P, _P, c_int, c_int, _P, c_i64, _P, c_int, _P, c_i64, _P,
                                _P, c_i64, _P, c_i64, _P, _P, _P],
    "ctr_p2p_alloc": [c_i64, _P],
    "ctr_p2p_free": [_P],
    "ctr_p2p_export": [_P, _P],
    "ctr_p2p_open": [_P, _P],
    "ctr_p2p_close": [_P],
    "ctr_p2p_barrier": [_P, _P, c_int, c_int, _P, _P],
    "ctr_shard_request": [_P, c_i64, c_i64, c_int, _P, _P, c_int, c_int, _P, _P, _P, _P, c_i64, _P, c_int, _P],
    "ctr_shard_serve": [c_int, c_int, c_int, _P, _P, c_i64, _P, _P, _P, _P, _P],
    "ctr_gather_fwd_exchanged": [_P, c_i64, c_i64, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P, c_int, _P,
                                 c_int, _P, _P, _P, c_i64, _P, _P, _P, c_int, c_int, _P, c_int, _P, _P, _P, _P, c_int, _P],
    "ctr_rowgrad_push": [c_i64, c_int, c_int, _P, _P, c_int, c_int, _P, c_i64, _P, c_int, _P, c_i64, _P,
                         _P, _P, _P, _P, c_i64, _P, _P],
    "ctr_lin_dense_wgrad": [_P, c_i64, c_i64, c_int, _P, _P, _P, _P],
    "ctr_dnn_layer_fwd": [_P, c_i64, _P, c_i64, c_i64, _P, _P, c_i64, c_i64, c_int, c_int, c_int, _P],
    "ctr_dnn_layer_bwd": [_P, c_i64, _P, c_i64, c_i64, _P, c_i64, _P, c_i64, _P, c_i64, c_int,
                          _P, c_i64, c_i64, _P, c_i64, c_int, c_int, c_int, _P],
    "ctr_dnn_layer_bwd_chain": [_P, c_i64, _P, c_i64, c_i64, _P, c_i64, _P, c_i64, _P, c_i64,
                                _P, c_i64, c_i64, _P, c_i64, c_int, c_int, c_int, c_int, c_int, _P],
    "ctr_sgemm": [c_i64, c_i64, c_i64, _P, c_i64, c_i64, _P, c_i64, c_i64, _P, c_i64, c_int, _P],
    "ctr_rowdot_fwd": [_P, c_i64, _P, c_i64, c_int, _P, c_int, _P],
    "ctr_rowdot_bwd": [_P, c_i64, _P, _P, c_i64, c_int, _P, c_i64, c_int, _P, _P],
    "ctr_predict_fwd": [_P, c_int, _P, c_i64, c_int, _P, _P, _P],
    "ctr_predict_bwd": [_P, _P, c_i64, c_int, _P, _P, _P],
    "ctr_cin_layer_fwd": [_P, c_i64, c_int, _P, c_i64, c_int, c_int, _P, _P, c_int, c_int, c_int,
                          _P, _P, c_i64, c_i64, _P],
    "ctr_cin_layer_bwd": [_P, c_i64, c_int, _P, c_i64, c_int, c_int, _P, c_int, c_int, c_int, c_int,
                          _P, _P, c_i64, _P, c_i64, _P, _P, _P, _P, c_i64, _P, c_i64, c_i64, _P],
    "ctr_cross_vector_fwd": [_P, c_i64, _P, _P, c_int, c_int, _P, c_i64, _P, c_i64, _P],
    "ctr_cross_vector_bwd": [_P, c_i64, _P, _P, c_int, c_int, _P, _P, c_i64, _P, c_i64, c_int,
                             _P, _P, c_i64, _P],
    "ctr_cross_matrix_layer_fwd": [_P, c_i64, _P, c_i64, _P, _P, c_int, _P, c_i64, _P, c_i64,
                                   c_i64, _P],
    "ctr_cross_matrix_layer_bwd": [_P, c_i64, _P, c_i64, _P, _P, c_i64, _P, c_i64, c_int, _P, c_i64,
                                   _P, c_i64, _P, _P, _P, c_i64, c_i64, _P],
    "ctr_cross_mix_fwd": [_P, c_i64, _P, c_i64, _P, c_i64, _P, _P, c_int, c_int, _P, c_i64, c_i64, _P],
    "ctr_cross_mix_bwd": [_P, c_i64, _P, c_i64, _P, _P, _P, c_i64, c_int, c_int, _P, _P, _P, c_i64,
                          _P, c_i64, _P],
    "ctr_senet_fwd": [_P, c_i64, c_int, c_int, _P, _P, c_int, _P, c_i64, c_i64, _P],
    "ctr_senet_bwd": [_P, c_i64, c_int, c_int, _P, _P, c_int, _P, c_i64, _P, c_i64, c_int, _P, _P,
                      c_i64, _P],
    "ctr_bilinear_fwd": [_P, c_i64, c_int, c_int, _P, c_int, _P, c_i64, c_i64, _P],
    "ctr_bilinear_bwd": [_P, c_i64, c_int, c_int, _P, c_int, _P, c_i64, _P, c_i64, _P, c_i64, _P],
    "ctr_sumsq_acc": [_P, c_i64, c_f32, _P, _P],
    "ctr_debug_set_buffer": [_P],
    "ctr_set_scratch": [_P, c_i64],
    "ctr_varlen_pool_fwd": [_P, c_i64, c_i64, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P,
                            c_i64, _P, c_int, _P],
    "ctr_varlen_pool_bwd": [_P, c_i64, c_i64, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P,
                            c_i64, _P, c_int, _P],
    "ctr_bipool_fwd": [_P, c_i64, c_int, c_int, _P, c_i64, c_i64, _P],
    "ctr_bipool_bwd": [_P, c_i64, c_int, c_int, _P, c_i64, _P, c_i64, c_i64, _P],
    "ctr_refine_fwd": [_P, _P, c_i64, _P, c_int, c_int, c_int, _P, _P, c_i64, _P, c_i64, _P],
    "ctr_refine_bwd": [_P, _P, c_i64, _P, c_int, c_int, c_int, _P, c_i64, _P, _P, _P, c_i64, _P, c_i64, _P],
    "ctr_afm_fwd": [_P, c_i64, c_int, c_int, c_int, _P, _P, _P, _P, c_i64, _P],
    "ctr_afm_bwd": [_P, c_i64, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_i64, _P, _P, _P, c_i64, _P],
    "ctr_fieldattn_fwd": [_P, _P, _P, _P, c_int, c_int, c_int, c_f32, _P, c_i64, _P],
    "ctr_fieldattn_bwd": [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_f32, _P, _P, _P, _P, c_i64, _P],
    "ctr_bce_sum_fwd": [_P, _P, c_i64, _P, _P],
    "ctr_bce_sum_bwd": [_P, _P, _P, c_i64, _P, _P],
    "ctr_prelu_fwd": [_P, _P, c_i64, _P, _P],
    "ctr_prelu_bwd": [_P, _P, _P, c_i64, _P, _P, _P],
    "ctr_rowopt_tick": [_P, _P],
    "ctr_rowopt_step": [c_int, c_i64, _P, _P, c_int, c_int, _P, c_i64, _P, _P, _P, _P, _P, c_f32, _P],
    "ctr_rowgrad_combine": [c_i64, c_int, c_int, _P, _P, _P, c_int, _P, _P, c_i64, _P, c_i64, _P],
}
SPECIAL = {
    "ctr_version": ([], c_int),
    "ctr_last_error": ([], ctypes.c_char_p),
    "ctr_unique_plan_hash_slots": ([c_i64], c_i64),
    "ctr_launch_count": ([], c_i64),
    "ctr_set_gemm_passes": ([c_int], c_int),
    "ctr_gemm_scratch_bytes": ([c_i64, c_i64, c_i64], c_i64),
    "ctr_dnn_wgrad_is_scratch_free": ([_P, c_i64, _P, c_i64, _P, c_i64, _P, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int], c_int),
}


class CtrLibraryError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def exported_symbols():
    return sorted(list(SIGNATURES) + list(SPECIAL))


def load():
    """Load libctr_b200.so once; raise loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CtrLibraryError(
                "libctr_b200.so not found at %s — build it with `python -m deepctr_torch_b200._build` "
                "(nvcc, sm_100a).  deepctr_torch_b200 has no CPU / torch fallback for the hot path." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        for name, (argtypes, restype) in SPECIAL.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        if lib.ctr_version() != ABI_VERSION:
            raise CtrLibraryError("libctr_b200.so ABI version %d, expected %d" % (lib.ctr_version(), ABI_VERSION))
        _lib = lib
    return _lib


_timing = None   # None, or {entry point: [(start_event, end_event), ...]} while profiling


def enable_timing(on=True):
    """Bracket every entry-point call with CUDA events on the current stream (bench/profiling)."""
    global _timing
    _timing = {} if on else None


def timing_summary():
    """{entry point: (calls, total_ms)} — synchronises."""
    import torch
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (_timing or {}).items()}


def launch_count():
    return int(load().ctr_launch_count())


def call(name, *args):
    """Invoke an int-returning entry point; non-zero return -> CtrLibraryError(message)."""
    lib = load()
    if _timing is not None:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(lib, name)(*args)
        b.record()
        _timing.setdefault(name, []).append((a, b))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.ctr_last_error()
        raise CtrLibraryError("%s returned %d: %s" % (name, rc, msg.decode() if msg else "?"))
