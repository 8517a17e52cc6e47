"""ctypes binding of the C-ABI library (include/tetgs_rast.h).

The product path has NO fallback: if libtetgs_rast.so is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtetgs_rast.so")

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)


class TgrParams(C.Structure):
    """Mirror of `tgr_params` (include/tetgs_rast.h). Pointers are passed as integers (c_void_p)."""

    _fields_ = [
        ("P", C.c_int32), ("D", C.c_int32), ("M", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32), ("extras", C.c_int32), ("accumulate", C.c_int32),
        ("depth_key_bits", C.c_int32), ("reserved_", C.c_int32),
        ("background", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p),
        ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p), ("image_buffer", C.c_void_p),
        ("geom_bytes", C.c_uint64), ("binning_bytes", C.c_uint64), ("image_bytes", C.c_uint64),
        ("out_color", C.c_void_p), ("radii", C.c_void_p), ("out_depth", C.c_void_p), ("out_alpha", C.c_void_p),
        ("dL_dout_color", C.c_void_p), ("dL_dout_depth", C.c_void_p), ("dL_dout_alpha", C.c_void_p),
        ("dL_dmeans2D", C.c_void_p), ("dL_dcolors", C.c_void_p), ("dL_dopacity", C.c_void_p),
        ("dL_dmeans3D", C.c_void_p), ("dL_dcov3D", C.c_void_p), ("dL_dsh", C.c_void_p),
        ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
        ("host_num_rendered", C.c_void_p),
    ]


class TgrBinding(C.Structure):
    """Mirror of `tgr_binding`."""

    _fields_ = [
        ("n_verts", C.c_int32), ("n_faces", C.c_int32),
        ("verts", C.c_void_p), ("vert_normals", C.c_void_p), ("faces", C.c_void_p), ("face_index", C.c_void_p),
        ("bary", C.c_void_p), ("delta", C.c_void_p), ("log_scales", C.c_void_p), ("raw_quats", C.c_void_p),
        ("opacity_logits", C.c_void_p),
        ("out_means3D", C.c_void_p), ("out_scales", C.c_void_p), ("out_rotations", C.c_void_p),
        ("out_opacities", C.c_void_p),
        ("dL_ddelta", C.c_void_p), ("dL_dlog_scales", C.c_void_p), ("dL_draw_quats", C.c_void_p),
        ("dL_dopacity_logits", C.c_void_p), ("dL_dverts", C.c_void_p),
        ("origins", C.c_void_p), ("normals", C.c_void_p), ("n_frozen", C.c_int32), ("reserved_", C.c_int32),
    ]


class TgrAdamGroup(C.Structure):
    """Mirror of `tgr_adam_group`."""

    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("count", C.c_uint64), ("lr", C.c_double), ("lr_alt", C.c_double), ("period", C.c_uint32), ("split", C.c_uint32),
    ]


# every symbol include/tetgs_rast.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "tgr_abi_version": (C.c_int, []),
    "tgr_last_error": (C.c_char_p, []),
    "tgr_geom_bytes": (C.c_uint64, [C.c_int32]),
    "tgr_image_bytes": (C.c_uint64, [C.c_int32, C.c_int32]),
    "tgr_binning_bytes": (C.c_uint64, [C.c_int32, C.c_uint64, C.c_int32, C.c_int32]),
    "tgr_binning_capacity": (C.c_uint64, [C.c_int32, C.c_uint64, C.c_int32, C.c_int32]),
    "tgr_forward_preprocess": (C.c_int, [C.POINTER(TgrParams), C.POINTER(TgrBinding), C.c_void_p]),
    "tgr_forward_render": (C.c_int, [C.POINTER(TgrParams), C.c_uint64, C.c_void_p]),
    "tgr_wait_num_rendered": (C.c_int, []),
    "tgr_backward": (C.c_int, [C.POINTER(TgrParams), C.POINTER(TgrBinding), C.c_uint64, C.c_void_p]),
    "tgr_forward_preprocess_batch": (C.c_int, [C.POINTER(TgrParams), C.c_int32, C.POINTER(TgrBinding), C.c_void_p]),
    "tgr_forward_depth_sort": (C.c_int, [C.POINTER(TgrParams), C.c_void_p]),
    "tgr_backward_blend": (C.c_int, [C.POINTER(TgrParams), C.c_uint64, C.c_void_p]),
    "tgr_forward_render_batch": (C.c_int, [C.POINTER(TgrParams), C.POINTER(C.c_uint64), C.c_int32, C.c_void_p]),
    "tgr_backward_blend_batch": (C.c_int, [C.POINTER(TgrParams), C.POINTER(C.c_uint64), C.c_int32, C.c_void_p]),
    "tgr_backward_preprocess_batch": (C.c_int, [C.POINTER(TgrParams), C.POINTER(C.c_uint64), C.c_int32,
                                                C.POINTER(TgrBinding), C.c_int32, C.c_int32, C.c_void_p]),
    "tgr_multimem_allreduce_f32": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p]),
    "tgr_multimem_allreduce_f32_capped": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tgr_read_header": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32 * 4), C.c_void_p]),
    "tgr_mark_visible": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_knn_bytes": (C.c_uint64, [C.c_int32]),
    "tgr_dist2": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tgr_export_binning": (C.c_int, [C.POINTER(TgrParams), C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "tgr_export_geom": (C.c_int, [C.POINTER(TgrParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_export_image_state": (C.c_int, [C.POINTER(TgrParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_kernel_launches": (C.c_uint64, []),
    "tgr_profile_enable": (C.c_int, [C.c_int]),
    "tgr_profile_collect": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "tgr_sort_temp_bytes": (C.c_uint64, [C.c_uint64]),
    "tgr_sort_pairs_u32_batch": (C.c_int, [C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_void_p]),
    "tgr_sort_pairs_u32": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_uint64, C.c_void_p]),
    "tgr_image_loss_bytes": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32]),
    "tgr_image_loss_forward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_void_p]),
    "tgr_image_loss_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_void_p]),
    "tgr_image_loss": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                 C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                 C.c_void_p]),
    "tgr_adam_step": (C.c_int, [C.POINTER(TgrAdamGroup), C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                C.c_float, C.c_void_p]),
    "tgr_mt_classify_bytes": (C.c_uint64, [C.c_int64]),
    "tgr_mt_edges_bytes": (C.c_uint64, [C.c_int64]),
    "tgr_mt_classify": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                  C.POINTER(C.c_uint32), C.c_void_p]),
    "tgr_mt_edges": (C.c_int, [C.c_int32, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                               C.POINTER(C.c_uint32), C.c_void_p]),
    "tgr_mt_emit": (C.c_int, [C.c_int64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_build_cameras": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
}

ABI_VERSION = 4
MAX_BATCH = 8
ADAM_MAX_GROUPS = 16
CAMERA_FLOATS = 40
NUM_STAGES = 8
STAGE_NAMES = ["preprocess", "depth_sort", "emit", "tile_sort", "ranges", "blend_fwd", "blend_bwd", "preprocess_bwd"]

_lib = None


class TgrError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libtetgs_rast.so (built by youreditableavatar_b200.build). Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TgrError(
                "libtetgs_rast.so not found at %s — build it with `python -m youreditableavatar_b200.build` "
                "(there is no CPU / PyTorch fallback for the rasterizer)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.tgr_abi_version() != ABI_VERSION:
            raise TgrError("ABI version mismatch")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().tgr_last_error()
        raise TgrError("%s failed (status %d): %s" % (what or "tgr call", rc, msg.decode() if msg else "?"))
