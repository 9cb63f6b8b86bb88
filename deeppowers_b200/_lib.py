"""ctypes loader for the in-tree C-ABI library (include/dpfhe.h).

There is no CPU fallback: if libdpfhe.so is missing it is built with nvcc, and if that is
impossible the import fails loudly.
"""
import ctypes as C
import os

from . import build as _build

_lib = None

SYMBOLS = {
    # name: (restype, argtypes)
    "dpfhe_last_error": (C.c_char_p, []),
    "dpfhe_version": (C.c_char_p, []),
    "dpfhe_context_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "dpfhe_context_destroy": (None, [C.c_void_p]),
    "dpfhe_get_modulus": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "dpfhe_get_psi": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "dpfhe_get_root_powers": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]),
    "dpfhe_context_device_bytes": (C.c_size_t, [C.c_void_p]),
    "dpfhe_ntt_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_ntt_inv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_poly_mul_pointwise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_poly_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_ct_tensor": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_keyswitch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_ct_mul_relin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_ct_mul_plain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_ct_mul_plain_acc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_rotate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_rotate_hoisted": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_ct_mul_plain_inner": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_mod_switch_down": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_keyswitch_hybrid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_ct_mul_relin_hybrid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_rotate_hybrid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_grouped_digits": (C.c_int, [C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]),
    "dpfhe_keyswitch_grouped": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_ct_mul_relin_grouped": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_rotate_grouped": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_rotate_hoisted_grouped": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64,
                                               C.c_void_p]),
    "dpfhe_mod_down_special": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
    "dpfhe_mod_down_special_host": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    "dpfhe_ct_mul_relin_grouped_host": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    "dpfhe_rotate_grouped_host": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    "dpfhe_ct_mul_relin_hybrid_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    "dpfhe_rotate_hybrid_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    "dpfhe_mod_switch_down_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    "dpfhe_fill_uniform": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_ntt_fwd_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_ntt_inv_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_ct_mul_relin_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_ct_mul_plain_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_rotate_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_context_trim": (C.c_int, [C.c_void_p]),
    "dpfhe_context_device": (C.c_int, [C.c_void_p]),
    "dpfhe_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "dpfhe_synchronize": (C.c_int, [C.c_void_p]),
    "dpfhe_galois_element": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
    "dpfhe_rotate_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_host_alloc_near": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t, C.POINTER(C.c_int)]),
    "dpfhe_device_numa_node": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "dpfhe_bind_thread_near": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "dpfhe_device_alloc": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t]),
    "dpfhe_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dpfhe_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p]),
    "dpfhe_ipc_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "dpfhe_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dpfhe_multi_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "dpfhe_multi_destroy": (None, [C.c_void_p]),
    "dpfhe_multi_device_count": (C.c_int, [C.c_void_p]),
    "dpfhe_multi_context": (C.c_void_p, [C.c_void_p, C.c_int]),
    "dpfhe_multi_shard": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "dpfhe_multi_ct_mul_relin_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_multi_ct_mul_relin_grouped_host": (C.c_int, [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]),
    "dpfhe_multi_rotate_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_multi_ct_mul_relin_gather": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_size_t]),
    "dpfhe_linear_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dpfhe_linear_destroy": (None, [C.c_void_p]),
    "dpfhe_linear_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dpfhe_linear_apply_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dpfhe_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "dpfhe_host_free": (C.c_int, [C.c_void_p]),
    "dpfhe_launch_count": (C.c_uint64, [C.c_void_p]),
    "dpfhe_debug_phase_cycles": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dpfhe_describe": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
}


class dpfhe_params(C.Structure):
    _fields_ = [("log_n", C.c_uint32), ("n_limbs", C.c_uint32), ("moduli", C.POINTER(C.c_uint64))]


def so_path():
    return _build.SO


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.SO
    if not os.path.exists(path):
        _build.build()          # raises if nvcc is unavailable: no fallback
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export what include/dpfhe.h declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
