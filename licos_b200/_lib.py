"""ctypes binding of liblicos_b200.so (the C ABI declared in include/licos_b200.h).

There is no fallback: if the shared library cannot be loaded (or built with nvcc), importing this
module raises, and every op of the package raises with it.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "liblicos_b200.so")

LICOS_OK = 0
ABI_VERSION = 3
ERR_NAMES = {
    -1: "LICOS_ERR_INVALID", -2: "LICOS_ERR_CUDA", -3: "LICOS_ERR_UNSUPPORTED", -4: "LICOS_ERR_NO_DEVICE",
    -5: "LICOS_ERR_DOMAIN", -6: "LICOS_ERR_NOMEM", -7: "LICOS_ERR_BUFFER",
}

LAYOUT_NCHW_F32 = 0
LAYOUT_NHWC_BF16 = 1
LAYOUT_NCHW_U8, LAYOUT_NCHW_U16, LAYOUT_NCHW_U16_Q8 = 2, 3, 4
CONV_5X5_S2 = 0
DECONV_5X5_S2 = 1
CONV_3X3_S1 = 2
CONV_1X1 = 3
EPI_NONE, EPI_GDN, EPI_IGDN, EPI_RELU = 0, 1, 2, 3
EB_MAX_LAYERS = 8
EB_FORM_PLAIN, EB_FORM_STABLE = 0, 1
EB_LUT_RADIUS = 128

c_int, c_i64, c_u64, c_f32, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p


class ConvArgs(ctypes.Structure):
    _fields_ = [
        ("kind", c_int), ("epilogue", c_int), ("batch", c_int), ("in_h", c_int), ("in_w", c_int),
        ("in_c", c_int), ("out_c", c_int), ("in_layout", c_int), ("out_layout", c_int),
        ("in_", c_vp), ("out", c_vp), ("weight", c_vp), ("bias", c_vp), ("beta", c_vp), ("gamma", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_i64), ("sm_count", c_int), ("int_max", c_int), ("pre_act", c_vp),
    ]


class EbRawPtrs(ctypes.Structure):
    """licos_eb_raw_params / licos_eb_raw_grads (same layout)."""
    _fields_ = [("matrix", c_vp * EB_MAX_LAYERS), ("bias", c_vp * EB_MAX_LAYERS), ("factor", c_vp * EB_MAX_LAYERS)]


class WgradArgs(ctypes.Structure):
    _fields_ = [
        ("kind", c_int), ("batch", c_int), ("h", c_int), ("w", c_int), ("big_h", c_int), ("big_w", c_int),
        ("small_c", c_int), ("big_c", c_int), ("small_t", c_vp), ("big_t", c_vp), ("out", c_vp),
        ("sm_count", c_int), ("reserved", c_int),
    ]


class EbParams(ctypes.Structure):
    _fields_ = [
        ("channels", c_int), ("n_layers", c_int), ("widths", c_int * (EB_MAX_LAYERS + 1)),
        ("params_per_channel", c_int), ("packed", c_vp), ("medians", c_vp), ("form", c_int),
        ("likelihood_bound", c_f32),
    ]


class EbFusedArgs(ctypes.Structure):
    _fields_ = [
        ("x", c_vp), ("batch", c_int), ("lut_ready", c_int), ("hw", c_i64), ("lut", c_vp), ("y_hat", c_vp), ("lik", c_vp),
        ("symbols", c_vp), ("symbols_i16", c_vp), ("y_hat_nhwc_bf16", c_vp), ("sum_ln", c_vp),
    ]


# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "licos_abi_version": (c_int, []),
    "licos_strerror": (ctypes.c_char_p, [c_int]),
    "licos_last_cuda_error": (c_int, []),
    "licos_device_ok": (c_int, [c_int]),
    "licos_pixel_scale_exact": (c_int, [c_int, c_int]),
    "licos_nchw_f32_to_nhwc_bf16": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_int, c_vp]),
    "licos_nhwc_bf16_to_nchw_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "licos_packed_weight_bytes": (c_i64, [c_int, c_int, c_int, c_int]),
    "licos_pack_conv_weight": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "licos_gdn_pack": (c_int, [c_vp, c_vp, c_int, c_f32, c_f32, c_f32, c_vp, c_vp, c_vp]),
    "licos_conv_workspace_bytes": (c_i64, [ctypes.POINTER(ConvArgs)]),
    "licos_conv_forward": (c_int, [ctypes.POINTER(ConvArgs), c_vp]),
    "licos_conv_wgrad": (c_int, [ctypes.POINTER(WgradArgs), c_vp]),
    "licos_gdn_backward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "licos_square_bf16": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "licos_gdn_bwd_mid": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "licos_gdn_bwd_out": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "licos_gdn_param_grad": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_f32, c_f32, c_vp, c_vp, c_vp]),
    "licos_relu_bwd": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "licos_colsum_bf16": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "licos_im2col5x5s2_kpad": (c_i64, [c_int]),
    "licos_im2col5x5s2": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "licos_conv_wgrad_image": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "licos_eb_lut_floats": (c_i64, [c_int]),
    "licos_eb_forward_eval": (c_int, [ctypes.POINTER(EbParams), c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "licos_eb_build_lut": (c_int, [ctypes.POINTER(EbParams), c_vp, c_vp]),
    "licos_eb_eval_fused": (c_int, [ctypes.POINTER(EbParams), ctypes.POINTER(EbFusedArgs), c_vp]),
    "licos_eb_forward_noise": (c_int, [ctypes.POINTER(EbParams), c_vp, c_vp, c_u64, c_int, c_i64, c_vp, c_vp, c_vp]),
    "licos_eb_backward": (c_int, [ctypes.POINTER(EbParams), c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_vp, c_vp]),
    "licos_eb_pack_params": (c_int, [ctypes.POINTER(EbRawPtrs), c_int, c_int, ctypes.POINTER(c_int), c_vp, c_vp]),
    "licos_eb_param_grads": (c_int, [ctypes.POINTER(EbRawPtrs), c_vp, c_int, c_int, ctypes.POINTER(c_int),
                                     ctypes.POINTER(EbRawPtrs), c_vp]),
    "licos_eb_aux_loss": (c_int, [c_vp, c_int, c_int, ctypes.POINTER(c_int), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "licos_eb_symbols": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp]),
    "licos_eb_dequantize": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp, c_vp]),
    "licos_gc_forward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_u64, c_i64, c_int, c_f32, c_f32, c_vp, c_vp, c_vp]),
    "licos_gc_backward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "licos_gc_build_indexes": (c_int, [c_vp, c_i64, c_vp, c_int, c_f32, c_vp, c_vp]),
    "licos_gc_symbols": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "licos_sum_log": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "licos_sum_sq_err": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "licos_scaled_reciprocal": (c_int, [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "licos_scaled_diff": (c_int, [c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "licos_pmf_to_quantized_cdf": (c_int, [c_vp, c_int, c_int, c_vp]),
    "licos_msssim_workspace_floats": (c_i64, [c_i64, c_int, c_int]),
    "licos_msssim_level": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp]),
    "licos_avgpool2": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "licos_raw_dn_to_unit": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "licos_rans_encode": (c_i64, [c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_i64]),
    "licos_rans_decode": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "licos_rans_encode_batch": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp,
                                        c_i64, c_vp, c_int]),
    "licos_rans_decode_batch": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_int, c_vp, c_vp,
                                        c_vp, c_int]),
    "licos_rans_encode_device": (c_int, [c_vp, c_vp, c_i64, c_int, c_i64, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                         c_i64, c_vp, c_vp]),
    "licos_rans_pack_device": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp]),
    "licos_rans_decode_device": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp,
                                         c_vp, c_vp]),
    "licos_weighted_sum2": (c_int, [c_vp, c_vp, c_f32, c_f32, c_i64, c_vp, c_vp]),
    "licos_scale_inplace": (c_int, [c_vp, c_f32, c_i64, c_vp]),
    "licos_nccl_version": (c_int, []),
    "licos_nccl_unique_id": (c_int, [c_vp]),
    "licos_nccl_comm_create": (c_int, [c_vp, c_int, c_int, ctypes.POINTER(c_vp)]),
    "licos_nccl_comm_destroy": (c_int, [c_vp]),
    "licos_nccl_weighted_allreduce": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        # in-tree build (nvcc cross-compiles without a GPU); raises if nvcc is missing or fails
        from .csrc.build import build
        build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.licos_abi_version() != ABI_VERSION:
        raise ImportError(f"liblicos_b200.so ABI version {lib.licos_abi_version()} != {ABI_VERSION}")
    return lib


lib = _load()


class LicosError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> int:
    """Translate a negative status into the Python exception the CompressAI call would raise."""
    if rc >= 0:
        return rc
    msg = lib.licos_strerror(int(rc)).decode()
    name = ERR_NAMES.get(int(rc), str(rc))
    if rc == -2:
        msg += f" [cudaError {lib.licos_last_cuda_error()}]"
    if rc in (-1, -5, -7):
        raise ValueError(f"{what}: {name}: {msg}")
    if rc == -3:
        raise NotImplementedError(f"{what}: {name}: {msg}")
    raise LicosError(f"{what}: {name}: {msg}")
