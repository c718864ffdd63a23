"""ctypes binding of libb200lp.so (the C ABI declared in include/b200lp.h).

PyTorch is used here only as the owner of device memory and of the current CUDA stream; every call passes raw
device pointers + sizes across the C ABI.  There is NO fallback: if the library is missing or the device is not
sm_100, calls raise.
"""
import ctypes
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent.parent
LIB_PATH = _PKG / "lib" / "libb200lp.so"
ABI_VERSION = 18


class B200lpError(RuntimeError):
    pass


class ConvArgs(Structure):
    _fields_ = [
        ("x", c_void_p), ("wp", c_void_p), ("out_scale", c_void_p), ("bias", c_void_p), ("residual", c_void_p),
        ("y", c_void_p), ("y_split", c_void_p),
        ("N", c_int32), ("H", c_int32), ("W", c_int32), ("Cin", c_int32), ("Cout", c_int32),
        ("ksize", c_int32), ("residual_mode", c_int32), ("relu", c_int32), ("round_tf32", c_int32),
        ("block_n", c_int32), ("precision", c_int32), ("stages", c_int32), ("ctas_per_sm", c_int32),
        ("splits", c_int32), ("variant", c_int32), ("a_stages", c_int32), ("grouped", c_int32), ("reserved0", c_int32),
        ("workspace", c_void_p),
        ("workspace_bytes", c_int64),
    ]


class WgradArgs(Structure):
    _fields_ = [
        ("x", c_void_p), ("dy", c_void_p), ("dw", c_void_p), ("workspace", c_void_p),
        ("workspace_bytes", c_int64),
        ("N", c_int32), ("H", c_int32), ("W", c_int32), ("Cin", c_int32), ("Cout", c_int32),
        ("ksize", c_int32), ("scale", c_float),
        ("kstep", c_int32), ("stages", c_int32), ("splits", c_int32), ("grouped", c_int32),
    ]


class SnItem(Structure):
    _fields_ = [
        ("w", c_void_p), ("u", c_void_p), ("v", c_void_p), ("snap_u", c_void_p), ("snap_v", c_void_p),
        ("scratch", c_void_p), ("inv_sigma", c_void_p),
        ("rows", c_int32), ("cols", c_int32), ("eps", c_float), ("reserved", c_int32),
    ]


_P = c_void_p
_I = c_int32
_L = c_int64
_F = c_float

# name -> (restype, argtypes); mirrors include/b200lp.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "b200lp_abi_version": (_I, []),
    "b200lp_last_error": (c_char_p, []),
    "b200lp_device_cc": (_I, []),
    "b200lp_launch_count": (_L, []),
    "b200lp_conv_fwd_workspace": (_L, [POINTER(ConvArgs)]),
    "b200lp_conv_fwd": (_I, [POINTER(ConvArgs), _P]),
    "b200lp_pack_conv_weight": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200lp_pack_conv_weight_multi": (_I, [_P, _P, _P, _I, _L, _P]),
    "b200lp_pack_conv_weight_tiles": (_I, [_P, _P, _P, _I, _P]),
    "b200lp_pack_gconv_weight": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200lp_sn_max_tensors": (_I, []),
    "b200lp_sn_scratch_floats": (_L, [_I, _I]),
    "b200lp_sn_sigma_multi": (_I, [POINTER(SnItem), _I, _I, _P]),
    "b200lp_sn_wgrad_fix_workspace": (_L, [_L]),
    "b200lp_sn_wgrad_fix": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "b200lp_conv_wgrad_workspace": (_L, [_I, _I, _I, _I, _I, _I]),
    "b200lp_conv_wgrad": (_I, [POINTER(WgradArgs), _P]),
    "b200lp_conv_wgrad_sn_acc_workspace": (_L, [_I, _I, _I, _I, _I, _I]),
    "b200lp_conv_wgrad_sn_acc": (_I, [POINTER(WgradArgs), _P, _P, _P, _P, _I, _P]),
    "b200lp_gconv3x3_wgrad_tc_workspace": (_L, [_I, _I, _I, _I]),
    "b200lp_gconv3x3_wgrad_tc": (_I, [POINTER(WgradArgs), _I, _P]),
    "b200lp_zero_stuff2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200lp_in_stats_workspace": (_L, [_I, _I, _I]),
    "b200lp_in_stats": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _F, _P]),
    "b200lp_adain_relu": (_I, [_P, _P, _P, _P, _P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_adain_relu_fused": (_I, [_P, _P, _P, _L, _P, _P, _P, _P, _P, _L, _P, _I, _I, _I, _I, _F, _I, _I, _P]),
    "b200lp_adain_relu_bwd_workspace": (_L, [_I, _I, _I]),
    "b200lp_adain_relu_bwd": (_I, [_P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _I, _P]),
    "b200lp_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200lp_nhwc_to_nchw": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200lp_relu_round": (_I, [_P, _P, _L, _P]),
    "b200lp_relu_bwd": (_I, [_P, _P, _P, _L, _P]),
    "b200lp_relu_bwd_fused": (_I, [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P]),
    "b200lp_avgpool2": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200lp_avgpool2_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200lp_upsample2_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200lp_l1_sum": (_I, [_P, _P, _P, _L, _F, _P]),
    "b200lp_l1_bwd": (_I, [_P, _P, _P, _F, _P, _L, _I, _P]),
    "b200lp_l1_relu_bwd": (_I, [_P, _P, _P, _F, _P, _P, _L, _P]),
    "b200lp_l1_sum_code": (_I, [_P, _P, _P, _P, _L, _F, _P]),
    "b200lp_l1_code_bwd": (_I, [_P, _P, _F, _P, _P, _L, _P]),
    "b200lp_l1_sum_code_pool": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "b200lp_l1_code_bwd_unpool": (_I, [_P, _P, _F, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_conv3x3_c3_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_conv3x3_c3_fwd_tc": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_conv3x3_c3_dgrad": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_conv3x3_c3_wgrad": (_I, [_P, _P, _P, _F, _I, _I, _I, _I, _P]),
    "b200lp_gen_tail_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_gen_tail_compose": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_gen_tail_bwd_act": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_gen_tail_bwd_data": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200lp_im2col3x3_c3": (_I, [_P, _P, _I, _I, _I, _P]),
    "b200lp_col2im3x3_c3": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "b200lp_gen_tail_bwd_weight": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_bias_grad": (_I, [_P, _P, _L, _I, _P]),
    "b200lp_bias_grad_acc": (_I, [_P, _P, _L, _I, _P]),
    "b200lp_copy_multi": (_I, [_P, _I, _P]),
    "b200lp_pw_conv_parts": (_I, [_L, _I]),
    "b200lp_pw_conv": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _L, _I, _I, _P]),
    "b200lp_pw_conv_workspace": (_L, [_L, _I, _I]),
    "b200lp_pw_conv_ws": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _L, _I, _I, _P, _L, _P]),
    "b200lp_dw_conv3x3_parts": (_I, [_I, _I, _I, _I]),
    "b200lp_dw_conv3x3": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200lp_mbv2_stem_parts": (_I, [_I, _I, _I]),
    "b200lp_mbv2_stem": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "b200lp_bn_finalize": (_I, [_P, _I, _L, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _I, _I, _P]),
    "b200lp_bn_apply": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _P]),
    "b200lp_bn_relu6_avgpool": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "b200lp_adam_ema_multi": (_I, [_P, _P, _P, _I, _L, _P, _F, _F, _F, _I, _I, _P]),
    "b200lp_ema_multi": (_I, [_P, _P, _P, _I, _L, _F, _P]),
    "b200lp_col_stats_parts": (_I, [_L]),
    "b200lp_col_stats": (_I, [_P, _P, _L, _I, _P]),
    "b200lp_bn_act": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _P]),
    "b200lp_bn_bwd_workspace": (_L, [_L, _I]),
    "b200lp_bn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _L, _L, _I, _I, _I, _I, _P]),
    "b200lp_gconv3x3_parts": (_I, [_I, _I, _I, _I, _I, _I]),
    "b200lp_gconv3x3_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_gconv3x3_dgrad": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_gconv3x3_wgrad_workspace": (_L, [_I, _I, _I, _I, _I, _I]),
    "b200lp_gconv3x3_wgrad": (_I, [_P, _P, _P, _P, _P, _I, _P, _L, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_im2col7x7_s2": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_maxpool3x3s2_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200lp_maxpool3x3s2_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_subsample2": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_scatter_add2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200lp_avgpool_fwd": (_I, [_P, _P, _I, _I, _I, _P]),
    "b200lp_avgpool_bwd": (_I, [_P, _P, _I, _I, _I, _P]),
    "b200lp_sgemm_strided_workspace": (_L, [_I, _I, _I]),
    "b200lp_sgemm_strided": (_I, [_P, _L, _L, _P, _L, _L, _P, _P, _P, _I, _I, _I, _I, _P, _L, _P]),
    "b200lp_dice_workspace": (_L, [_I, _I]),
    "b200lp_dice_fwd": (_I, [_P, _P, _F, _P, _P, _P, _L, _I, _I, _I, _P]),
    "b200lp_dice_bwd": (_I, [_P, _P, _P, _P, _F, _P, _I, _I, _I, _P]),
    "b200lp_adversarial_fwd": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "b200lp_adversarial_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "b200lp_crop_bilinear_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_crop_bilinear_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200lp_disc_head_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "b200lp_disc_head_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "b200lp_transpose2d": (_I, [_P, _P, _I, _I, _P]),
    "b200lp_pw_wgrad_workspace": (_L, [_L, _I, _I]),
    "b200lp_pw_wgrad": (_I, [_P, _P, _P, _P, _I, _P, _I, _P, _L, _L, _I, _I, _P]),
    "b200lp_dw_dgrad": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "b200lp_dw_wgrad_workspace": (_L, [_I, _I, _I, _I, _I]),
    "b200lp_dw_wgrad": (_I, [_P, _P, _P, _P, _P, _I, _P, _L, _I, _I, _I, _I, _I, _P]),
    "b200lp_mbv2_stem_wgrad_workspace": (_L, [_I, _I, _I]),
    "b200lp_mbv2_stem_wgrad": (_I, [_P, _P, _P, _I, _P, _L, _I, _I, _I, _P]),
}

_lib = None
c_void_p = c_void_p   # re-exported for callers that pass raw pointers


def load(build_if_missing=True):
    """dlopen libb200lp.so, declare every signature, verify the ABI version.  Raises B200lpError if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing:
            import importlib.util
            spec = importlib.util.spec_from_file_location("b200lp_build_ext", _PKG / "build_ext.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        if not LIB_PATH.exists():
            raise B200lpError(f"{LIB_PATH} not found: run `python __graft_entry__.py` (build) first")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise B200lpError(f"libb200lp.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    v = lib.b200lp_abi_version()
    if v != ABI_VERSION:
        raise B200lpError(f"libb200lp.so ABI version {v} != binding version {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def last_error():
    return load().b200lp_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise B200lpError(f"{what} failed (code {rc}): {last_error()}")


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=torch.float32):
    """Raw device pointer of a contiguous CUDA tensor of the given dtype (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise B200lpError("b200lp kernels need CUDA tensors (there is no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise B200lpError(f"expected contiguous {dtype} tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return c_void_p(t.data_ptr())


_device_ok = {}


def require_device():
    """Fail loudly unless a CUDA device of compute capability 10.x is current (checked once per device: the query is
    not something to repeat inside a CUDA-graph capture)."""
    if not torch.cuda.is_available():
        raise B200lpError("no CUDA device: the b200lp hot path has no CPU fallback")
    dev = torch.cuda.current_device()
    cc = _device_ok.get(dev)
    if cc is None:
        cc = load().b200lp_device_cc()
        if cc < 100 or cc >= 110:
            raise B200lpError(f"device compute capability {cc} is not sm_100 (B200): kernels are sm_100a-only")
        _device_ok[dev] = cc
    return cc
