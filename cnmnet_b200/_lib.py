"""ctypes binding of libcnmnet_b200.so (C ABI in include/cnmnet_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
Pointers come from ``torch.Tensor.data_ptr()``; PyTorch is only the allocator / stream owner.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcnmnet_b200.so")

CNM_BF16, CNM_TF32, CNM_BF16X3, CNM_F16F8 = 0, 1, 2, 3
VOLUME_NCHW_F32, VOLUME_ACT = 0, 1


class ConvSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("apron", C.c_int32),
                ("x_off", C.c_int32), ("y_off", C.c_int32), ("pitch_w", C.c_int32), ("pitch_hw", C.c_int32)]


class ConvStep(C.Structure):
    _fields_ = [("src", C.c_int32), ("sub", C.c_int32), ("dx", C.c_int32), ("dy", C.c_int32), ("c0", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class ConvDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
                ("c_out_total", C.c_int32), ("c_out_offset", C.c_int32), ("stride", C.c_int32), ("relu", C.c_int32),
                ("num_srcs", C.c_int32), ("num_steps", C.c_int32), ("dtype", C.c_int32), ("reserved", C.c_int32),
                ("acc_exp", C.c_int32), ("ksplit", C.c_int32), ("cluster", C.c_int32), ("tile_h", C.c_int32), ("fold", C.c_int32), ("balance", C.c_int32), ("head_w", C.c_void_p),
                ("head_partial", C.c_void_p), ("splitk_ws", C.c_void_p), ("side_disp", C.c_void_p), ("side_w", C.c_void_p)]


class SweepFuse(C.Structure):
    _fields_ = [("ref", C.c_void_p), ("packed_src", C.c_void_p), ("hk", C.c_void_p), ("scratch", C.c_void_p),
                ("scratch_bytes", C.c_int64), ("B", C.c_int32), ("D", C.c_int32), ("vol_src", C.c_int32),
                ("reserved", C.c_int32), ("idepth_min", C.c_double), ("idepth_max", C.c_double)]


_P, _I, _D, _F, _L = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_int64

# name -> (restype, argtypes); must list every function declared in include/cnmnet_b200.h
SIGNATURES = {
    "cnm_last_error": (C.c_char_p, []),
    "cnm_version": (C.c_char_p, []),
    "cnm_launch_count": (_L, []),
    "cnm_set_range_check": (None, [_I]),
    "cnm_set_pdl": (None, [_I]),
    "cnm_range_overflow_count": (_L, [_I]),
    "cnm_camera_setup": (_I, [_P, _P, _P, _I, _I, _P]),
    "cnm_packed_source_bytes": (_L, [_I, _I, _I]),
    "cnm_pack_source": (_I, [_P, _P, _I, _I, _I, _P]),
    "cnm_cost_volume": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _D, _D, _I, _I, _P]),
    "cnm_plane_sweep": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _D, _D, _I, _I, _P]),
    "cnm_depth2normal": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cnm_intrinsics_inverse": (_I, [_P, _P, _I, _P]),
    "cnm_loss_workspace_bytes": (_L, []),
    "cnm_idepth_loss": (_I, [_P, _P, _P, _P, _P, _L, _I, _P]),
    "cnm_idepth_loss_234": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "cnm_surface_normal_loss": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "cnm_depth_metrics_workspace_bytes": (_L, [_I, _I, _I]),
    "cnm_depth_metrics": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _F, _P]),
    "cnm_plane_regularize_workspace_bytes": (_L, [_I, _I, _I, _I]),
    "cnm_plane_regularize": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cnm_reciprocal": (_I, [_P, _P, _L, _P]),
    "cnm_nchw_to_act": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "cnm_act_to_nchw": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "cnm_pack_row_taps": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "cnm_pack_row_taps3": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "cnm_pack_disp_taps": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "cnm_refine_aggregate": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "cnm_refine_aggregate_nchw": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "cnm_upsample2x": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "cnm_head3x3": (_I, [_P, _P, _F, _F, _P, _P, _I, _I, _I, _I, _I, _P]),
    "cnm_head_partial": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "cnm_head_finish": (_I, [_P, _I, _F, _F, _P, _P, _I, _I, _I, _P]),
    "cnm_conv2d_tc": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvSrc), _P, _P, _P, _P, _P]),
    "cnm_conv_balance_ws_bytes": (_L, []),
    "cnm_debug_conv_timeline": (_I, [C.POINTER(C.c_longlong)]),
    "cnm_conv2d_simt": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvSrc), _P, _P, _P, _P, _P]),
    "cnm_conv2d_halo": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvSrc), C.POINTER(C.c_int32), _P, _I, _P, _P, _P, _P]),
    "cnm_conv2d_halo2": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvSrc), C.POINTER(C.c_int32), _P, _I, _P, _P, _P, _P]),
    "cnm_sweep_scratch_bytes": (_L, [_I, _I]),
    "cnm_conv2d_halo2_sweep": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvSrc), C.POINTER(C.c_int32), _P, _I, _P, _P, _P,
                                    C.POINTER(SweepFuse), _P]),
    "cnm_inverse_warp": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
}

_lib = None


def load():
    """Loads the shared library (once) and types its entry points.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("cnmnet_b200: %s not found -- build it with `python -m cnmnet_b200.build` "
                           "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cnm_last_error().decode("utf-8", "replace")
        raise RuntimeError("cnmnet_b200 %s failed (%d): %s" % (what, rc, msg))


def launch_count():
    return int(load().cnm_launch_count())


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("cnmnet_b200: expected a CUDA tensor (no CPU path exists)")
    if not t.is_contiguous():
        raise RuntimeError("cnmnet_b200: expected a contiguous tensor")
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
