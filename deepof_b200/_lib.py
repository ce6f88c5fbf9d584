"""ctypes binding of libdeepof_b200.so (the C ABI declared in include/deepof_b200.h).

The product path has NO fallback: if the shared library is missing or a CUDA device is not
usable, importing a compute entry point raises.  (The CPU oracle in oracle/ is test
infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libdeepof_b200.so"


class DeepOFError(RuntimeError):
    pass


class ConvGeom(C.Structure):
    _fields_ = [("B", C.c_int), ("ih", C.c_int), ("iw", C.c_int), ("ci", C.c_int),
                ("oh", C.c_int), ("ow", C.c_int), ("co", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad_t", C.c_int), ("pad_l", C.c_int)]


class PackJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("taps", C.c_int), ("ci", C.c_int), ("co", C.c_int), ("contract_ci", C.c_int)]


class LossScale(C.Structure):
    _fields_ = [("flow", C.c_void_p), ("src", C.c_void_p), ("tgt", C.c_void_p),
                ("recon", C.c_void_p), ("dflow", C.c_void_p), ("loss4", C.c_void_p),
                ("B", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("flow_scale", C.c_float),
                ("epsilon", C.c_float), ("alpha_c", C.c_float), ("alpha_s", C.c_float), ("lambda_smooth", C.c_float),
                ("g_charb", C.c_float), ("g_u", C.c_float), ("g_v", C.c_float),
                ("variant", C.c_int), ("edge_w", C.c_void_p)]


class _StencilEntry(C.Structure):
    _fields_ = [("dy", C.c_int), ("dx", C.c_int), ("cin", C.c_int), ("cout", C.c_int), ("w", C.c_float)]


class FlowStencil(C.Structure):
    _fields_ = [("n", C.c_int), ("e", _StencilEntry * 64)]


_P, _I, _F, _LL = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_G = C.POINTER(ConvGeom)

# name -> (restype, argtypes); every symbol include/deepof_b200.h declares
SIGNATURES = {
    "dofb_version": (_I, []),
    "dofb_last_error": (C.c_char_p, []),
    "dofb_launch_count": (_LL, []),
    "dofb_reset_launch_count": (None, []),
    "dofb_maxpool2_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "dofb_maxpool2_bwd": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "dofb_preprocess": (_I, [_P, _P, C.POINTER(C.c_float), _F, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_void_p),
                             C.POINTER(C.c_void_p), _P]),
    "dofb_preprocess_bf16": (_I, [_P, _P, C.POINTER(C.c_float), _F, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_void_p), _P]),
    "dofb_preprocess_u8": (_I, [_P, _P, C.POINTER(C.c_float), _F, _I, _I, _I, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, C.POINTER(C.c_void_p),
                                C.POINTER(C.c_void_p), _P]),
    "dofb_conv1_fwd": (_I, [_G, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P]),
    "dofb_conv_fwd_bf16": (_I, [_G, _P, _I, _P, _P, _P, _P, _I, _I, _P]),
    "dofb_conv_dgrad_bf16": (_I, [_G, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P]),
    "dofb_conv_wgrad_bf16": (_I, [_G, _P, _I, _P, _I, _P, _P]),
    "dofb_cast_bf16": (_I, [_P, _I, _P, _I, _LL, _I, _P]),
    "dofb_conv1_wgrad": (_I, [_G, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "dofb_conv1_fwd_bf16": (_I, [_G, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P]),
    "dofb_conv1_wgrad_bf16": (_I, [_G, _P, _I, _I, _I, _I, _P, _I, _P, _P]),
    "dofb_warp_loss_workspace_bytes": (C.c_size_t, [_I, C.POINTER(LossScale)]),
    "dofb_warp_loss": (_I, [_I, C.POINTER(LossScale), _P, C.c_size_t, _P]),
    "dofb_edge_weights_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "dofb_edge_weights": (_I, [_P, _I, _I, _I, _P, _P, C.c_size_t, _P]),
    "dofb_warp_loss_multi_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "dofb_warp_loss_multi": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _F, _F, _F, _F, C.POINTER(FlowStencil), _P, C.c_size_t, _P]),
    "dofb_decode_ppm": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _P]),
    "dofb_decode_flo": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "dofb_eval_flow_aee_sum": (_I, [_P, _I, _I, _I, _P, _I, _I, _F, _F, _F, _P, _P]),
    "dofb_conv_fwd": (_I, [_G, _P, _I, _P, _P, _P, _I, _I, _I, _P]),
    "dofb_conv_dgrad": (_I, [_G, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P]),
    "dofb_conv_wgrad": (_I, [_G, _P, _I, _P, _I, _P, _P, _I, _P]),
    "dofb_conv_wgrad_tbias": (_I, [_G, _P, _I, _P, _I, _P, _P, _I, _P]),
    "dofb_elu_bwd": (_I, [_P, _I, _P, _I, _LL, _I, _P, _P, _P]),
    "dofb_elu_bwd_shadow": (_I, [_P, _I, _P, _I, _LL, _I, _P, _P, _P]),
    "dofb_elu_bwd_shadow16": (_I, [_P, _I, _P, _I, _LL, _I, _P, _P, _P]),
    "dofb_invalidate_weight_cache": (None, []),
    "dofb_enable_weight_cache": (None, [_I]),
    "dofb_enable_cta_pairs": (None, [_I]),
    "dofb_enable_halo_tiles": (None, [_I]),
    "dofb_enable_phase_in_n": (None, [_I]),
    "dofb_enable_split_k": (None, [_I]),
    "dofb_enable_wgrad_npack": (None, [_I]),
    "dofb_pack_weights_batch": (_I, [_P, _I, _I, _P]),
    "dofb_head_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "dofb_head_dgrad": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    "dofb_head_wgrad": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P]),
    "dofb_uppr_fwd": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "dofb_uppr_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "dofb_head_wz_pack": (_I, [_I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), _P]),
    "dofb_head_wgrad_bf16": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "dofb_head_tapsum": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "dofb_head_dpr9": (_I, [_P, _I, _I, _I, _P, _I, _P, _P]),
    "dofb_head_dgrad_elu_bf16": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _P]),
    "dofb_adam": (_I, [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _F, _P]),
    "dofb_epe_sum": (_I, [_P, _P, _LL, _P, _P]),
    "dofb_corr_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    "dofb_corr_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _I, _I, _P]),
    "dofb_corr_fwd_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    "dofb_corr_bwd_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P]),
}

_lib = None


def load(path: os.PathLike | None = None) -> C.CDLL:
    """dlopen the library and attach argtypes.  Does not touch the GPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise DeepOFError(f"{p} is missing: build it with `python -m deepof_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.dofb_version() != 100:
        raise DeepOFError(f"libdeepof_b200 version {lib.dofb_version()} does not match the Python binding (100)")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, lib: C.CDLL | None = None) -> None:
    if rc != 0:
        msg = (lib or load()).dofb_last_error()
        raise DeepOFError(msg.decode() if msg else f"libdeepof_b200 call failed (rc={rc})")
