"""ctypes binding of libbluetangle_cuda.so (include/bluetangle_cuda.h).

Stands in for the Julia ``ccall`` stubs of julia/BlueTangleCUDA.jl in this container (no Julia here).
There is deliberately no fallback: if the shared library is missing or no CUDA device is visible, calls fail.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbluetangle_cuda.so")


class BTError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbluetangle_cuda error {code}: {msg}")
        self.code = code


class bt_c64(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class bt_gate(C.Structure):
    _fields_ = [("nq", C.c_int32), ("qubit", C.c_int32), ("target", C.c_int32), ("control", C.c_int32), ("m", bt_c64 * 16)]


class bt_dm_op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nq", C.c_int32), ("qubit", C.c_int32), ("target", C.c_int32), ("control", C.c_int32), ("nK", C.c_int32), ("mats", C.c_void_p)]


GATE_DTYPE = np.dtype([("nq", "<i4"), ("qubit", "<i4"), ("target", "<i4"), ("control", "<i4"), ("m", "<c16", (16,))])
assert GATE_DTYPE.itemsize == C.sizeof(bt_gate)

BARRIER_FN = C.CFUNCTYPE(None, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int)

_vp, _i, _i64, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
_pd = C.POINTER(C.c_double)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)

# name -> argtypes; every function returns int except bt_last_error.  This table is also what
# tests/test_abi.py checks against include/bluetangle_cuda.h.
PROTOTYPES = {
    "bt_version": [],
    "bt_device_count": [C.POINTER(_i)],
    "bt_set_device": [_i],
    "bt_set_strict": [_i],
    "bt_fp64_peak": [_pd, _pd, _i],
    "bt_fusion_plan_host": [_i, _vp, _u64, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _pd, C.POINTER(_i), _i, C.POINTER(_i)],
    "bt_fusion_stats": [C.POINTER(_u64), C.POINTER(_u64)],
    "bt_fusion_flops": [_pd],
    "bt_jit_stats": [C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), _pd],
    "bt_jit_selftest": [C.c_char_p, _u64],
    "bt_jit_debug_only": [_i, _i],
    "bt_jit_verify_stats": [C.POINTER(_u64), C.POINTER(_u64)],
    "bt_jit_config": [_pi32, _pi32, _pi32, _pi32],
    "bt_jit_wait": [C.POINTER(_u64)],
    "bt_jit_selftest_workers": [_i],
    "bt_jit_cache_info": [C.POINTER(_u64), C.c_char_p, _u64],
    "bt_sv_create": [_i, _i64, C.POINTER(_vp)],
    "bt_sv_destroy": [_vp],
    "bt_pool_release": [],
    "bt_sv_n_qubits": [_vp, C.POINTER(_i)],
    "bt_sv_set_basis": [_vp, _u64],
    "bt_sv_set_plus": [_vp],
    "bt_sv_upload": [_vp, _vp, _u64],
    "bt_sv_download": [_vp, _vp, _u64],
    "bt_sv_copy": [_vp, _vp],
    "bt_sv_sync": [_vp],
    "bt_sv_timer_start": [_vp],
    "bt_sv_timer_stop": [_vp, C.POINTER(C.c_float)],
    "bt_sv_launch_count": [_vp, C.POINTER(_u64)],
    "bt_sv_profile_enable": [_vp, _i],
    "bt_sv_profile_read": [_vp, C.POINTER(_u64), _pd],
    "bt_sv_apply_1q": [_vp, _i, _vp, _i],
    "bt_sv_apply_2q": [_vp, _i, _i, _vp, _i],
    "bt_sv_apply_3q": [_vp, _i, _vp],
    "bt_sv_apply_circuit": [_vp, _vp, _u64, _i],
    "bt_sv_apply_1q_if": [_vp, _i, _vp, _i, _i],
    "bt_sv_apply_2q_if": [_vp, _i, _i, _vp, _i, _i],
    "bt_sv_rdm1": [_vp, _i, _vp],
    "bt_sv_rdm2": [_vp, _i, _i, _vp],
    "bt_sv_rdm3": [_vp, _i, _vp],
    "bt_sv_rdm": [_vp, _i, C.POINTER(_i), _vp],
    "bt_sv_schmidt_spectrum": [_vp, _i, _pd, C.POINTER(_i)],
    "bt_jacobi_pairs_host": [_i, _i, C.POINTER(_i)],
    "bt_sv_expect_op": [_vp, _i, _i, _i, _i, _vp, _pd],
    "bt_sv_norm2": [_vp, _pd],
    "bt_sv_inner": [_vp, _vp, _vp],
    "bt_sv_normalize": [_vp],
    "bt_sv_probs": [_vp, _pd],
    "bt_sv_measure_z": [_vp, _i, _pd, _pi32, _pd, _i],
    "bt_sv_measure_z_multi": [_vp, _i, C.POINTER(_i), _pd, _pi32, C.POINTER(_i)],
    "bt_sv_measure_log": [_vp, _i],
    "bt_sv_measure_log_read": [_vp, _pi32, _u64, C.POINTER(_u64)],
    "bt_sv_outcomes": [_vp, _pi32],
    "bt_sv_set_mask": [_vp, _pi32],
    "bt_sv_kraus": [_vp, _i, _i, _i, _vp, _i, _pd, _pi32],
    "bt_sv_kraus_probs": [_vp, _i, _i, _i, _vp, _i, _pd],
    "bt_sv_expect_pauli": [_vp, C.c_char_p, _pd],
    "bt_sv_expect_1q_all": [_vp, _vp, _pd],
    "bt_sv_expect_product": [_vp, _i, C.POINTER(_i), _vp, _pd],
    "bt_sv_expect_matrix2q": [_vp, _i, _i, _vp, _pd],
    "bt_sv_expect_pauli_sum": [_vp, _i, C.c_char_p, _pd, _pd],
    "bt_sv_sample": [_vp, _pd, _u64, _pi64],
    "bt_sv_sample_batched": [_vp, _pd, _u64, _pi64],
    "bt_dm_create": [_i, C.POINTER(_vp)],
    "bt_dm_destroy": [_vp],
    "bt_dm_n_qubits": [_vp, C.POINTER(_i)],
    "bt_dm_from_sv": [_vp, _vp],
    "bt_dm_upload": [_vp, _vp, _u64],
    "bt_dm_download": [_vp, _vp, _u64],
    "bt_dm_sync": [_vp],
    "bt_dm_timer_start": [_vp],
    "bt_dm_timer_stop": [_vp, C.POINTER(C.c_float)],
    "bt_dm_launch_count": [_vp, C.POINTER(_u64)],
    "bt_dm_apply_1q": [_vp, _i, _vp, _i],
    "bt_dm_apply_2q": [_vp, _i, _i, _vp, _i],
    "bt_dm_kraus": [_vp, _i, _i, _i, _vp, _i],
    "bt_dm_dephase": [_vp, _i],
    "bt_dm_apply_circuit": [_vp, _vp, _u64, _i],
    "bt_dm_apply_ops": [_vp, _vp, _u64, _i],
    "bt_dm_diag": [_vp, _pd],
    "bt_dm_trace": [_vp, _vp],
    "bt_dm_expect_pauli": [_vp, C.c_char_p, _pd],
    "bt_dm_expect_1q_all": [_vp, _vp, _pd],
    "bt_dm_expect_product": [_vp, _i, C.POINTER(_i), _vp, _pd],
    "bt_dm_sample": [_vp, _pd, _u64, _pi64],
    "bt_dm_rdm": [_vp, _i, C.POINTER(_i), _vp],
    "bt_dm_expect_op": [_vp, _i, _i, _i, _i, _vp, _pd],
    "bt_dm_bipartition_spectrum": [_vp, _i, _pd, C.POINTER(_i)],
    "bt_dm_fidelity": [_vp, _vp, _pd],
    "bt_sv_create_shard": [_i, _i, _i, C.POINTER(_vp)],
    "bt_sv_ipc_export": [_vp, _vp],
    "bt_sv_ipc_attach": [_vp, _vp],
    "bt_sv_attach_local_peers": [C.POINTER(_vp), _i],
    "bt_sv_set_barrier": [_vp, BARRIER_FN, _vp],
    "bt_sv_remap": [_vp, C.POINTER(_i)],
    "bt_group_apply_circuit": [C.POINTER(_vp), _i, _vp, _u64, _i],
    "bt_plan_circuit_host": [_i, _i, _vp, _u64, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i],
    "bt_sv_layout": [_vp, C.POINTER(_i)],
    "bt_sv_remap_stats": [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(C.c_float)],
    "bt_remap_walk_host": [_i, _i, _i, C.POINTER(_i), C.POINTER(_i), _u64, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)],
    "bt_sv_remap_log": [_vp, _i, C.POINTER(C.c_float), C.POINTER(_i)],
    "bt_sv_set_allreduce": [_vp, ALLREDUCE_FN, _vp],
}

_lib = None


def load(path: str | None = None) -> C.CDLL:
    """dlopen the library and attach prototypes.  Raises if it is not built -- the product never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    p = path or os.environ.get("BLUETANGLE_CUDA_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise BTError(-2, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    lib = C.CDLL(p)
    lib.bt_last_error.restype = C.c_char_p
    lib.bt_last_error.argtypes = []
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise BTError(rc, load().bt_last_error().decode("utf-8", "replace"))


def cmat(m, dim: int) -> np.ndarray:
    """Column-major (Julia layout) complex128 copy of a dim x dim matrix."""
    a = np.asarray(m, dtype=np.complex128)
    if a.shape != (dim, dim):
        raise ValueError(f"size of matrix {a.shape} not compatible with {dim}x{dim}")
    return np.asfortranarray(a)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pdouble(a: np.ndarray):
    return a.ctypes.data_as(_pd)
