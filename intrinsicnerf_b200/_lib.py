"""ctypes binding of libinrf.so (include/inrf.h).  The library is the product: if it is
missing or fails to load, every op raises - there is no Python/CPU fallback."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# INRF_LIB: developer override (e.g. a -DTC2_PROF profiling build of the same sources)
LIB_PATH = os.environ.get("INRF_LIB") or os.path.join(_HERE, "csrc", "libinrf.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "inrf.h")

PREC_TC, PREC_FP32 = 0, 1
NET_OBJECT, NET_SSR = 0, 1
RAW_BASE, REC_BASE = 11, 13

_lib = None

p, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float


class RenderCfg(C.Structure):
    _fields_ = [("variant", C.c_int32), ("n_classes", C.c_int32), ("n_samples", C.c_int32),
                ("n_importance", C.c_int32), ("lindisp", C.c_int32), ("white_bkgd", C.c_int32),
                ("endpoint_feat", C.c_int32), ("precision", C.c_int32), ("pe_scalar_factor", C.c_float),
                ("reserved", C.c_int32 * 7)]


class Camera(C.Structure):
    """InrfCamera (include/inrf.h)."""
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("c2w", C.c_float * 12), ("near_", C.c_float), ("far_", C.c_float), ("convention", C.c_int32),
                ("euclidean", C.c_int32), ("reserved", C.c_int32 * 4)]


class FramePlanes(C.Structure):
    """InrfFramePlanes (include/inrf.h): nullable output planes of inrf_frame_finish."""
    _fields_ = [(n, C.c_void_p) for n in ("rgb8", "albedo8", "shading8", "residual8", "label8", "vis_label8", "entropy8",
                                          "entropy", "disp16", "depth_mm16", "labels64", "sample_pixels", "sample_labels")] \
        + [("reserved", C.c_void_p * 3)]


_SIGS = {
    "inrf_last_error_string": (C.c_char_p, []),
    "inrf_version": (i32, []),
    "inrf_poll_status": (i32, []),
    "inrf_launch_count": (i64, []),
    "inrf_rng_epoch_bump": (i32, [p]),
    "inrf_rng_epoch_reset": (i32, [p]),
    "inrf_flat_param_count": (i64, [i32, i32]),
    "inrf_packed_bytes": (i64, [i32, i32]),
    "inrf_pack_weights": (i32, [p, i32, i32, p, i64, p]),
    "inrf_embed": (i32, [p, i64, i32, f32, p, p]),
    "inrf_mlp_fwd": (i32, [p, i32, i32, i32, f32, p, p, i64, p, i32, p]),
    "inrf_mlp_fwd_embedded": (i32, [p, i32, i32, i32, p, i64, p, i32, p]),
    "inrf_mlp_fwd_rays": (i32, [p, i32, i32, i32, f32, p, p, i64, i32, p, i32, p]),
    "inrf_stash_floats_per_row": (i64, []),
    "inrf_mlp_fwd_train": (i32, [p, i32, i32, i32, f32, p, p, p, p, i32, p, i64, p, p, p]),
    "inrf_mlp_bwd": (i32, [p, i32, i32, i32, f32, p, p, p, p, i32, p, i64, p, p, p, p, p]),
    "inrf_mlp_stash_img_bytes": (i64, [i64]),
    "inrf_mlp_bwd_tc_workspace_bytes": (i64, [i32, i32, i64]),
    "inrf_mlp_fwd_train_tc": (i32, [p, i32, i32, i32, f32, p, p, p, p, i32, p, i64, p, p, p]),
    "inrf_mlp_bwd_tc": (i32, [p, p, i32, i32, i32, i64, p, p, p, p, i64, p, p]),
    "inrf_raw2outputs": (i32, [p, p, p, i32, p, i64, i32, i32, i32, i32, p, p, p]),
    "inrf_raw2outputs_bwd": (i32, [p, p, p, i32, p, i64, i32, i32, i32, i32, p, p, p, p]),
    "inrf_sample_pdf": (i32, [p, p, i32, p, p, i64, i32, i32, p, p, p, p]),
    "inrf_invert_cdf": (i32, [p, p, p, i64, i32, i32, p, p, p]),
    "inrf_merge_sorted": (i32, [p, p, i64, i32, i32, p, p, p]),
    "inrf_coarse_z": (i32, [p, p, p, i64, i32, i32, p, p]),
    "inrf_coarse_z_rng": (i32, [p, p, C.c_uint64, i64, i32, i32, p, p]),
    "inrf_sample_pdf_rng": (i32, [p, p, i32, C.c_uint64, i64, i32, i32, p, p]),
    "inrf_raw2outputs_rng": (i32, [p, p, p, i32, f32, C.c_uint64, i32, i64, i32, i32, i32, i32, p, p, p]),
    "inrf_raw2outputs_bwd_rng": (i32, [p, p, p, i32, f32, C.c_uint64, i32, i64, i32, i32, i32, i32, p, p, p, p]),
    "inrf_get_rays": (i32, [i32, i32, f32, f32, f32, f32, C.POINTER(C.c_float), f32, f32, p, p]),
    "inrf_intrinsic_loss_fwd": (i32, [p, i32, p, i32, p, i32, p, i32, p, p, p, i64, i32, p, p]),
    "inrf_intrinsic_loss_bwd": (i32, [p, i32, p, i32, p, i32, p, i32, p, p, p, i64, i32, p, p, p, p, p, p]),
    "inrf_rays_from_pixels": (i32, [p, i64, i32, i32, f32, f32, f32, f32, C.POINTER(C.c_float), i32, i32, f32, f32, p, p]),
    "inrf_render_workspace_bytes": (i64, [C.POINTER(RenderCfg), i64]),
    "inrf_render_fwd": (i32, [p, i64, p, p, C.POINTER(RenderCfg)] + [p] * 13 + [p, i64, p]),
    "inrf_render_fwd_camera": (i32, [C.POINTER(Camera), i64, i64, p, p, C.POINTER(RenderCfg)] + [p] * 7 + [i64, p]),
    "inrf_mapping_color": (i32, [p, i64, f32, p, p]),
    "inrf_nearest_anchor": (i32, [p, i64, p, i64, i32, f32, p, p]),
    "inrf_dest_color": (i32, [p, i64, p, p, i64, p, i64, f32, p, p, p]),
    "inrf_choose_anchors": (i32, [p, p, i64, p, p, p, p, p]),
    "inrf_meanshift_seeds": (i32, [p, i64, p, i64, f32, i32, p, p, p, p]),
    "inrf_kth_neighbor_dist": (i32, [p, i64, p, i64, i32, p, p]),
    "inrf_frame_finish": (i32, [p, i32, i32, i32, i32, f32, p, i32, C.POINTER(FramePlanes), p]),
    "inrf_edit_recompose": (i32, [p, p, i64, i32, p, p, p]),
}


def declared_symbols():
    """Every function name include/inrf.h declares (used by the CPU test-suite)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(inrf_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m intrinsicnerf_b200.build` "
                "(there is no fallback implementation)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class InrfError(RuntimeError):
    pass


class InrfRangeError(InrfError):
    """INRF_ERANGE: a weight, activation or gradient left the fp16 range of the tensor-core path."""


E_RANGE = -5


def check(rc):
    if rc < 0:
        cls = InrfRangeError if rc == E_RANGE else InrfError
        raise cls(f"libinrf error {rc}: {lib().inrf_last_error_string().decode()}")
    return rc
