"""ctypes binding of libsgdm_b200.so (the C ABI declared in include/sgdm_b200.h).

There is NO fallback: if the shared library is missing, or a compute entry point is
called without a CUDA device, this module raises.  Build the library with
`python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SGDM_LIB selects another build of the same library (tests: the -DSGDM_OPERAND_BF16 variant)
LIB_PATH = os.environ.get("SGDM_LIB") or os.path.join(_HERE, "libsgdm_b200.so")

MAX_DIMS = 8


class SgdmConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("image_size", C.c_int32),
        ("in_channels", C.c_int32),
        ("out_channels", C.c_int32),
        ("model_channels", C.c_int32),
        ("num_res_blocks", C.c_int32),
        ("n_channel_mult", C.c_int32),
        ("channel_mult", C.c_int32 * 8),
        ("n_attention_resolutions", C.c_int32),
        ("attention_resolutions", C.c_int32 * 8),
        ("num_heads", C.c_int32),
        ("resblock_updown", C.c_int32),
        ("cond_dim", C.c_int32),
        ("layout_dim", C.c_int32),
        ("context_dim", C.c_int32),
        ("cond_token_num", C.c_int32),
        ("precision", C.c_int32),
        ("use_cls_token_as_pooled", C.c_int32),
    ]


PRECISIONS = {"fp16": 0, "fp16x3": 1}  # sgdm_config.precision
KIND_UNET_FAST = 0
KIND_UNETCA_FAST = 1
SCALE_TYPES = {"imagen": 0, "cfg": 1}

_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol declared in include/sgdm_b200.h
PROTOTYPES = {
    "sgdm_last_error": (C.c_char_p, []),
    "sgdm_version": (C.c_char_p, []),
    "sgdm_operand_dtype": (C.c_char_p, []),
    "sgdm_create": (_i, [C.POINTER(SgdmConfig), C.POINTER(_vp)]),
    "sgdm_destroy": (_i, [_vp]),
    "sgdm_param_count": (_i, [_vp]),
    "sgdm_param_name": (C.c_char_p, [_vp, _i]),
    "sgdm_param_shape": (_i, [_vp, _i, C.POINTER(_i64), C.POINTER(_i)]),
    "sgdm_load_param": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _vp]),
    "sgdm_params_missing": (_i, [_vp]),
    "sgdm_set_timestep_freqs": (_i, [_vp, _vp, _i]),
    "sgdm_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "sgdm_forward_guided": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_vp), C.POINTER(_vp)]),
    "sgdm_mix": (_i, [_vp, _vp, _vp, _d, _vp, _i, _vp, _i, _i64]),
    "sgdm_ddim_step": (_i, [_vp, _vp, _vp, _d, _vp, _i, C.POINTER(_f), _i, _vp, _vp, _vp, _vp, _vp, _i, _i64]),
    "sgdm_ddpm_step": (_i, [_vp, _vp, _vp, _d, _vp, _i, C.POINTER(_f), _i, _vp, _vp, _vp, _vp, _i, _i64]),
    "sgdm_ddim_step_ex": (_i, [_vp, _vp, _vp, _d, _vp, _i, C.POINTER(_f), _i, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "sgdm_ddpm_step_ex": (_i, [_vp, _vp, _vp, _d, _vp, _i, C.POINTER(_f), _i, _vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "sgdm_dyn_threshold": (_i, [_vp, _i, _vp, _vp, _d, _vp, _i, C.POINTER(_f), _vp, _d, _vp, _vp, _i, _i64]),
    "sgdm_lincomb": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_f), _f, _vp, _i64]),
    "sgdm_lincomb_scaled": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_f), _f, _vp, _i64]),
    "sgdm_pndm_transfer": (_i, [_vp, _vp, _vp, _f, _f, _f, _vp, _i64]),
    "sgdm_to_uint8": (_i, [_vp, _vp, _vp, _i64]),
    "sgdm_set_graph_mode": (_i, [_vp, _i]),
    "sgdm_set_share_prefix": (_i, [_vp, _i]),
    "sgdm_fingerprint": (_i, [_vp, _vp, _vp, _i, _vp]),
    "sgdm_set_profiling": (_i, [_vp, _i]),
    "sgdm_profile_count": (_i, [_vp]),
    "sgdm_profile_get": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_d), C.POINTER(_d), C.POINTER(_d)]),
    "sgdm_profile_executed_flops": (_i, [_vp, _i, C.POINTER(_d)]),
    "sgdm_launch_count": (_i64, []),
    "sgdm_debug_set_conv_pair": (_i, [_i]),
    "sgdm_debug_set_conv_timing": (_i, [_vp]),
    "sgdm_debug_set_conv_halo": (_i, [_i]),
    "sgdm_debug_set_conv_k32": (_i, [_i]),
    "sgdm_debug_set_conv_astat": (_i, [_i]),
    "sgdm_debug_set_attn_tc": (_i, [_i]),
    "sgdm_k_conv": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i]),
    "sgdm_k_conv_stats": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i,
                                _vp, _i, _vp, _vp, _i]),
    "sgdm_k_pack_weight": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    "sgdm_k_groupnorm": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "sgdm_k_groupnorm_fused": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _i, _vp, _vp,
                                     _vp]),
    "sgdm_k_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i]),
    "sgdm_k_layernorm_stats": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i]),
    "sgdm_k_layernorm_split3": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i]),
    "sgdm_k_attention": (_i, [_vp, _vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _i, _vp, _vp, _i, _vp, _i64, _i, _i, _i, _i, _f]),
    "sgdm_k_linear_f32": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i]),
    "sgdm_k_linear_f32_splitk": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _i]),
    "sgdm_k_conv_up2": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i]),
    "sgdm_k_conv_head_hfold": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "sgdm_k_cast": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i]),
    "sgdm_k_quantile_abs": (_i, [_vp, _vp, _i, _i64, _d, _vp]),
}

_lib = None


class SgdmError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SgdmError(
                f"{LIB_PATH} is missing: the CUDA extension is required (there is no CPU or eager fallback). "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise SgdmError(lib().sgdm_last_error().decode())


def ptr(t):
    """Device pointer of a (contiguous) torch tensor, or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    return t.data_ptr()


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise SgdmError(f"{what} is on {t.device}: sgdm_b200 runs on CUDA (sm_100a) only and has no CPU fallback")


def current_stream(device=None):
    import torch

    return torch.cuda.current_stream(device).cuda_stream
