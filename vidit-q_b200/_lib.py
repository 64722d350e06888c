"""ctypes binding of libviditq_b200.so (the C ABI declared in include/viditq_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libviditq_b200.so")

VQ_OK = 0
VQ_EPI_BIAS, VQ_EPI_GELU_TANH, VQ_EPI_GATE_RESIDUAL = 0, 1, 2
VQ_STATUS_EPS_DEGENERATE = 1
_ERR = {-1: "VQ_ERR_ARG", -2: "VQ_ERR_DRIVER", -3: "VQ_ERR_TMAP", -4: "VQ_ERR_LAUNCH", -5: "VQ_ERR_UNSUPPORTED"}

EXPORTS = ["vq_version", "vq_num_sms", "vq_prep_weight", "vq_act_quant", "vq_act_quant_static", "vq_gelu_act_quant_static", "vq_ln_modulate_act_quant_static", "vq_add_act_quant", "vq_gelu_act_quant", "vq_act_quant_heads", "vq_ln_modulate_act_quant", "vq_gemm_w8a8",
           "vq_col_absmax", "vq_row_pack", "vq_linear_w8a8", "vq_linear_workspace_bytes", "vq_linear_launch_count", "vq_linear_set_fused_policy", "vq_pack_u4", "vq_linear_w4a8",
           "vq_attn_temporal", "vq_attn_temporal_quant", "vq_attn_cross", "vq_attn_spatial", "vq_attn_i8_workspace_bytes", "vq_attn_i8_quantise", "vq_attn_i8_attend", "vq_attn_spatial_i8", "vq_cfg_ddim_step", "vq_patch_embed", "vq_status_read"]

_lib = None


class VqError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VqError(f"{LIB_PATH} not built: run `python -c \"import __graft_entry__ as g; g.build()\"` "
                      "(nvcc, sm_100a). viditq_b200 has no CPU fallback.")
    import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.vq_version.restype = i32
    L.vq_num_sms.restype = i32
    L.vq_prep_weight.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
    L.vq_act_quant.argtypes = [vp, i32, i32, i32, i64, i64, vp, i32, vp, vp, vp, vp, vp, vp]
    L.vq_add_act_quant.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp]
    L.vq_act_quant_static.argtypes = [vp, i32, i32, i64, vp, vp, i32, vp, i32, vp, vp, vp]
    L.vq_gelu_act_quant_static.argtypes = [vp, i32, i32, i64, vp, vp, i32, vp, i32, vp, vp, vp]
    L.vq_ln_modulate_act_quant_static.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp]
    L.vq_gelu_act_quant.argtypes = [vp, i32, i32, i32, i64, i64, vp, i32, vp, vp, vp, vp, vp, vp]
    L.vq_act_quant_heads.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.vq_ln_modulate_act_quant.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.vq_gemm_w8a8.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp]
    L.vq_col_absmax.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.vq_row_pack.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i64, i32, vp]
    L.vq_linear_workspace_bytes.argtypes = [i32, i32, i32]
    L.vq_linear_launch_count.argtypes = [i32, i32, i32]
    L.vq_linear_set_fused_policy.argtypes = [i32, i64]
    L.vq_linear_w8a8.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, vp, i32, vp, i32, vp, i32, vp, vp,
                                 vp, i64, vp, vp]
    L.vq_pack_u4.argtypes = [vp, i32, i32, vp, vp]
    L.vq_linear_w4a8.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, vp, i32, vp, i32, vp, i32, vp, vp,
                                 vp, vp]
    f32 = ctypes.c_float
    L.vq_attn_temporal.argtypes = [vp, vp, i32, i32, i32, i32, i32, f32, vp]
    L.vq_attn_spatial.argtypes = [vp, vp, i32, i32, i32, i32, f32, vp]
    L.vq_attn_i8_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.vq_attn_i8_quantise.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.vq_attn_i8_attend.argtypes = [vp, vp, i32, i32, i32, i32, f32, vp]
    L.vq_attn_spatial_i8.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, vp]
    L.vq_attn_temporal_quant.argtypes = [vp, i32, i32, i32, i32, i32, f32, vp, i32, vp, vp, vp, vp, vp, vp]
    L.vq_attn_cross.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, f32, vp]
    L.vq_cfg_ddim_step.argtypes = [vp, vp, vp, vp, f32, ctypes.c_double, i32, i32, i32, i64, vp, vp]
    L.vq_patch_embed.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    L.vq_status_read.argtypes = [vp, vp, vp]
    for name in EXPORTS:
        getattr(L, name).restype = i32
    L.vq_linear_workspace_bytes.restype = i64
    L.vq_attn_i8_workspace_bytes.restype = i64
    _lib = L
    return L


def check(rc, what):
    if rc != VQ_OK:
        raise VqError(f"{what} failed: {_ERR.get(rc, rc)}")
