"""ctypes binding of libedgegan_b200.so (the C ABI declared in include/edgegan_b200.h).

There is no fallback: if the shared library is missing the import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libedgegan_b200.so")

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class ConvShape(C.Structure):
    """eg_conv_shape (include/edgegan_b200.h)."""
    _fields_ = [(n, C.c_int) for n in
                ("N", "H", "W", "Ci", "OH", "OW", "Co", "KH", "KW", "stride", "pad_t", "pad_l")]


_csp = C.POINTER(ConvShape)


class FilterDesc(C.Structure):
    """eg_filter_desc (include/edgegan_b200.h)"""
    _fields_ = [("w", C.c_void_p), ("taps", C.c_int), ("Ci", C.c_int), ("Co", C.c_int)]

class SnDesc(C.Structure):
    """eg_sn_desc (include/edgegan_b200.h)"""
    _fields_ = [(n, C.c_void_p) for n in ("W", "u", "Wbar", "ws", "G", "gW", "Wa", "Wi", "Ga", "Gi")] + \
               [(n, C.c_int) for n in ("K", "C", "cin", "hd")]

# name -> argtypes ; every function returns int (0 = ok) unless listed in _RESTYPE
SIGNATURES = {
    "eg_abi_version": [],
    "eg_device_info": [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
    "eg_set_default_algo": [i32],
    "eg_get_default_algo": [],
    "eg_debug_set": [i32, i32],
    "eg_norm_debug": [i32],
    "eg_kernel_launches": [],
    "eg_filter_set_create": [vp, i32, i32, C.POINTER(i64)],
    "eg_filter_set_prepare": [i64, vp],
    "eg_filter_set_destroy": [i64],
    "eg_filter_set_hits": [],
    "eg_crc32c": [vp, i64, C.c_uint],
    "eg_conv2d_algo_for": [_csp, i32, i32],
    "eg_conv2d_fwd": [_csp, vp, vp, vp, vp, i32, vp],
    "eg_conv2d_bwd_data": [_csp, vp, vp, vp, vp, i32, vp],
    "eg_conv2d_bwd_weight": [_csp, vp, vp, vp, i32, i32, vp],
    "eg_conv2d_fwd_ex": [_csp, vp, vp, vp, vp, i32, i32, vp, i32, vp],
    "eg_conv2d_bwd_data_ex": [_csp, vp, vp, vp, vp, i32, i32, vp, i32, vp],
    "eg_bias_grad": [vp, i64, i32, vp, i32, vp],
    "eg_instnorm_fwd": [vp, vp, vp, i32, i32, i32, f32, i32, vp],
    "eg_instnorm_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, vp],
    "eg_instnorm_bwd2": [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, vp],
    "eg_act_fwd": [vp, vp, i64, i32, vp],
    "eg_act_bwd": [vp, vp, vp, i64, i32, vp],
    "eg_bn_stats": [vp, vp, i32, i32, vp],
    "eg_bn_apply": [vp, vp, f32, vp, vp, vp, i32, i32, f32, i32, vp],
    "eg_bn_bwd_reduce": [vp, vp, f32, vp, vp, vp, vp, i32, i32, f32, i32, vp],
    "eg_bn_bwd_apply": [vp, vp, f32, vp, vp, vp, vp, vp, i32, i32, f32, i32, vp],
    "eg_rowdot_fwd": [vp, vp, vp, vp, i32, i32, vp],
    "eg_rowdot_bwd_input": [vp, vp, vp, i32, i32, vp],
    "eg_rowdot_bwd_weight": [vp, vp, vp, vp, i32, i32, i32, vp],
    "eg_bicubic_up2_fwd": [vp, vp, i32, i32, i32, i32, vp],
    "eg_bicubic_up2_bwd": [vp, vp, i32, i32, i32, i32, vp],
    "eg_copy2d": [vp, i64, vp, i64, i64, i64, vp],
    "eg_fill": [vp, i64, f32, vp],
    "eg_u8_lut_f32": [vp, vp, vp, i64, vp],
    "eg_axpby": [vp, vp, i64, f32, f32, vp],
    "eg_gp_interpolate": [vp, vp, vp, vp, i32, i64, vp],
    "eg_gp_seed": [vp, vp, i32, vp],
    "eg_gp_penalty": [vp, vp, vp, vp, i32, i64, f32, f32, vp],
    "eg_gp_seed_bwd": [vp, vp, vp, i32, vp],
    "eg_sum_scaled": [vp, i64, f32, vp, i32, vp],
    "eg_reflect_pad_fwd": [vp, vp, i32, i32, i32, i32, i32, vp],
    "eg_reflect_pad_bwd": [vp, vp, i32, i32, i32, i32, i32, vp],
    "eg_addrelu_pool2_fwd": [vp, vp, vp, i32, i32, i32, i32, vp],
    "eg_addrelu_pool2_bwd": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "eg_relu_globalmean_fwd": [vp, vp, i32, i32, i32, vp],
    "eg_relu_globalmean_bwd": [vp, vp, vp, i32, i32, i32, vp],
    "eg_reparam_fwd": [vp, vp, f32, vp, vp, i64, vp],
    "eg_zl1_loss_bwd": [vp, vp, f32, vp, vp, i32, i32, i32, f32, f32, vp, vp, vp, vp],
    "eg_prelu_fwd": [vp, vp, vp, i64, vp],
    "eg_prelu_bwd": [vp, vp, vp, vp, vp, i64, i32, vp],
    "eg_prelu_fwd2": [vp, vp, vp, vp, vp, i64, vp],
    "eg_prelu_bwd_ex": [vp, vp, vp, vp, vp, i64, i32, i32, vp],
    "eg_mru_gate_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "eg_mru_gate_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "eg_minmax_fwd": [vp, vp, vp, i32, i32, i32, vp],
    "eg_minmax_bwd": [vp, vp, vp, vp, i32, i32, i32, vp],
    "eg_fma3": [vp, vp, vp, vp, i64, vp],
    "eg_mul": [vp, vp, vp, i64, vp],
    "eg_add_pool2_fwd": [vp, vp, vp, i32, i32, i32, i32, vp],
    "eg_add_pool2_prelu_fwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "eg_pool2_bwd": [vp, vp, i32, i32, i32, i32, i32, vp],
    "eg_globalmean_fwd": [vp, vp, i32, i32, i32, vp],
    "eg_globalmean_bwd": [vp, vp, i32, i32, i32, vp],
    "eg_spectral_norm_ws_floats": [i32, i32],
    "eg_spectral_norm_fwd": [vp, vp, vp, vp, i32, i32, vp],
    "eg_spectral_norm_bwd": [vp, vp, vp, vp, vp, i32, i32, vp],
    "eg_spectral_norm_set_create": [vp, i32, C.POINTER(i64)],
    "eg_spectral_norm_set_fwd": [i64, vp],
    "eg_spectral_norm_set_bwd": [i64, vp],
    "eg_spectral_norm_set_destroy": [i64],
    "eg_softmax_ce_bwd": [vp, vp, i32, i32, i32, i32, i32, f32, f32, vp, vp, vp],
    "eg_onehot_concat": [vp, i32, i32, i32, vp, vp],
    "eg_rmsprop": [vp, vp, vp, i64, f32, f32, f32, vp],
}
_RESTYPE = {"eg_last_error": C.c_char_p, "eg_kernel_launches": C.c_longlong, "eg_filter_set_hits": C.c_longlong, "eg_crc32c": C.c_uint}

_lib = None


def load():
    """Return the loaded library (ctypes.CDLL); raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C edgegan_b200/csrc`).  edgegan_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.eg_last_error.restype = C.c_char_p
    lib.eg_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library mismatch
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


def exported_symbols():
    return ["eg_last_error"] + list(SIGNATURES)


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().eg_last_error().decode(errors="replace")
        raise RuntimeError(f"edgegan_b200 {what} failed (rc={rc}): {msg}")
