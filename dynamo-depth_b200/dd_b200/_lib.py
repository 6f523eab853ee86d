"""ctypes binding of libdynamo_b200.so (C ABI declared in include/dynamo_b200.h).

The library is the product: if it cannot be loaded, every op raises -- there is no CPU or
PyTorch fallback on this path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DD_B200_LIB selects another build of the same library (kernel A/B experiments); the default is the in-tree build
LIB_PATH = os.environ.get("DD_B200_LIB") or os.path.join(HERE, "libdynamo_b200.so")

DD_MAX_SCALES = 4
DD_MAX_FRAMES = 2
DD_NSUM = 8
DD_FLAG_CMPFLOW = 1
DD_FLAG_MOTMASK = 2
DD_FLAG_AUTOMASK = 4
DD_SUM_PHOTO = 0
DD_SUM_CONSIST0 = 1
DD_SUM_MAG0 = 3
DD_SUM_IDENT = 5

FP = C.c_void_p  # raw device pointer


class WarpDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("num_scales", C.c_int32), ("num_frames", C.c_int32), ("flags", C.c_int32),
        ("min_depth", C.c_float), ("max_depth", C.c_float),
        ("ssim_weight", C.c_float), ("mask_disp_thrd", C.c_float),
        ("target", FP),
        ("source", FP * DD_MAX_FRAMES),
        ("K", FP), ("inv_K", FP),
        ("T", FP * DD_MAX_FRAMES),
        ("ts", FP * DD_MAX_FRAMES),
        ("scale", C.c_int32 * DD_MAX_SCALES),
        ("disp", FP * DD_MAX_SCALES),
        ("flow", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
        ("mask", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
        ("noise", FP * DD_MAX_SCALES),
    ]


class WarpAux(C.Structure):
    _fields_ = [
        ("warped", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
        ("sample", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
        ("depth", FP * DD_MAX_SCALES),
        ("ident_sel", FP * DD_MAX_SCALES),
        ("resid", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
        ("independ", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
        ("mag", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
    ]


class WarpGrads(C.Structure):
    _fields_ = [
        ("disp", FP * DD_MAX_SCALES),
        ("T", FP * DD_MAX_FRAMES),
        ("flow", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
        ("mask", (FP * DD_MAX_FRAMES) * DD_MAX_SCALES),
    ]


DD_MAX_SMOOTH_TASKS = 24


class SmoothTask(C.Structure):
    _fields_ = [
        ("inp", FP), ("img", FP), ("grad_inp", FP),
        ("B", C.c_int32), ("C", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("mean_normalise", C.c_int32),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cout", C.c_int32),
        ("ksize", C.c_int32), ("pad_mode", C.c_int32), ("act", C.c_int32), ("up0", C.c_int32),
        ("C0", C.c_int32), ("C1", C.c_int32),
        ("x0", FP), ("x1", FP), ("weight", FP), ("bias", FP), ("residual", FP),
    ]


PAD_ZERO, PAD_REFLECT = 0, 1
UP_NONE, UP_NEAREST2, UP_BILINEAR2 = 0, 1, 2
ACT_NONE, ACT_ELU, ACT_SIGMOID, ACT_RELU = 0, 1, 2, 3

# symbol -> (restype, argtypes); tests check that the .so exports exactly the header's entry points
SIGNATURES = {
    "dd_last_error": (C.c_char_p, []),
    "dd_version": (C.c_int, []),
    "dd_device_sm_count": (C.c_int, []),
    "dd_launch_count": (C.c_longlong, []),
    "dd_warp_photo_workspace_bytes": (C.c_size_t, [C.POINTER(WarpDesc)]),
    "dd_warp_photo_fwd": (C.c_int, [C.POINTER(WarpDesc), C.POINTER(WarpAux), FP, FP, C.c_size_t, FP]),
    "dd_warp_photo_bwd": (C.c_int, [C.POINTER(WarpDesc), FP, C.POINTER(WarpAux), C.POINTER(WarpGrads), FP, C.c_size_t, FP]),
    "dd_smooth_workspace_bytes": (C.c_size_t, [C.POINTER(SmoothTask), C.c_int]),
    "dd_smooth_fwd": (C.c_int, [C.POINTER(SmoothTask), C.c_int, FP, FP, C.c_size_t, FP]),
    "dd_smooth_bwd": (C.c_int, [C.POINTER(SmoothTask), C.c_int, FP, FP, C.c_size_t, FP]),
    "dd_msparsity_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dd_msparsity_fwd": (C.c_int, [FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, C.c_size_t, FP]),
    "dd_msparsity_bwd": (C.c_int, [FP, FP, FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_conv_workspace_bytes": (C.c_size_t, [C.POINTER(ConvDesc)]),
    "dd_conv_fwd": (C.c_int, [C.POINTER(ConvDesc), FP, FP, C.c_size_t, FP]),
    "dd_conv_bwd": (C.c_int, [C.POINTER(ConvDesc), FP, FP, FP, FP, FP, FP, FP, C.c_size_t, FP]),
    "dd_resize_bilinear_fwd": (C.c_int, [FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_resize_bilinear_bwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_pyramid_half_fwd": (C.c_int, [FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_backproject_fwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_backproject_bwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_project_fwd": (C.c_int, [FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, FP]),
    "dd_project_bwd": (C.c_int, [FP, FP, FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, FP]),
    "dd_ssim_fwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_ssim_bwd": (C.c_int, [FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_pose_mean_fwd": (C.c_int, [FP, C.c_int, C.c_int, C.c_float, FP, FP]),
    "dd_pose_mean_bwd": (C.c_int, [FP, C.c_int, C.c_int, C.c_float, FP, FP]),
    "dd_pose_matrix_fwd": (C.c_int, [FP, FP, C.c_int, C.c_int, FP, FP]),
    "dd_pose_matrix_bwd": (C.c_int, [FP, FP, FP, C.c_int, C.c_int, FP, FP, FP]),
    "dd_ground_score": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, FP, FP]),
    "dd_nchw_to_nhwc": (C.c_int, [FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_nhwc_to_nchw": (C.c_int, [FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_block_tail_fwd": (C.c_int, [FP, FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_block_tail_bwd": (C.c_int, [FP, FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, FP]),
    "dd_linear_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dd_linear_fwd": (C.c_int, [FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, C.c_size_t, FP]),
    "dd_linear_bwd": (C.c_int, [FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, FP, FP, C.c_size_t, FP]),
    "dd_layernorm_workspace_bytes": (C.c_size_t, [C.c_int]),
    "dd_layernorm_fwd": (C.c_int, [FP, C.c_longlong, C.c_int, FP, FP, C.c_float, FP, FP, FP, FP]),
    "dd_layernorm_bwd": (C.c_int, [FP, FP, C.c_longlong, C.c_int, FP, FP, FP, FP, FP, FP, FP, C.c_size_t, FP]),
    "dd_dwconv3x3_workspace_bytes": (C.c_size_t, [C.c_int]),
    "dd_dwconv3x3_fwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_dwconv3x3_wgrad": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP, C.c_size_t, FP]),
    "dd_maxpool3x3s2_nhwc_fwd": (C.c_int, [FP, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP, FP]),
    "dd_maxpool3x3s2_nhwc_bwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP]),
    "dd_xca_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dd_xca_fwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP, FP, FP, FP, FP, C.c_size_t, FP]),
    "dd_xca_bwd": (C.c_int, [FP, FP, FP, FP, FP, FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, FP, FP, FP, C.c_size_t, FP]),
    "dd_bn_nhwc_workspace_bytes": (C.c_size_t, [C.c_int]),
    "dd_bn_act_nhwc_fwd": (C.c_int, [FP, FP, C.c_longlong, C.c_int, FP, FP, C.c_float, C.c_float, C.c_int, FP, FP, FP, FP, FP, FP,
                                     C.c_size_t, FP]),
    "dd_bn_act_nhwc_bwd": (C.c_int, [FP, FP, FP, C.c_longlong, C.c_int, FP, FP, FP, FP, C.c_int, FP, FP, FP, FP, FP, C.c_size_t, FP]),
    "dd_bn_workspace_bytes": (C.c_size_t, [C.c_int]),
    "dd_bn_gelu_fwd": (C.c_int, [FP, C.c_int, C.c_int, C.c_int, FP, FP, C.c_float, C.c_float, C.c_int, FP, FP, FP, FP, FP, FP,
                                 C.c_size_t, FP]),
    "dd_bn_gelu_bwd": (C.c_int, [FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, FP, FP, C.c_int, FP, FP, FP, FP, C.c_size_t, FP]),
}

_lib = None


class DynamoB200Error(RuntimeError):
    pass


def load():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DynamoB200Error(
            f"{LIB_PATH} is missing: build it with `python dynamo-depth_b200/build.py` "
            "(the CUDA extension is mandatory, there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dd_last_error()
        raise DynamoB200Error(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a CUDA fp32 contiguous tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DynamoB200Error("dynamo_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype.is_floating_point and t.dtype != __import__("torch").float32:
        raise DynamoB200Error(f"dynamo_b200 ops need fp32 tensors, got {t.dtype}")
    if not t.is_contiguous():
        raise DynamoB200Error("dynamo_b200 ops need contiguous tensors")
    return t.data_ptr()
