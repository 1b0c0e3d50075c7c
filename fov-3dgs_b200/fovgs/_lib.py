"""ctypes binding of libfovgs.so (the C-ABI declared in include/fovgs.h).

The library is hand-written CUDA for sm_100a; there is NO fallback: if the shared object is missing or a
call fails, a RuntimeError is raised (the reference raises RuntimeError from AT_ERROR / CHECK_CUDA the same way,
FOV/rasterize_points.cu:64-66, FOV/cuda_rasterizer/auxiliary.h:298-305).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FOVGS_LIB_PATH: load another build of the same library (A/B measurements of compile-time parameters)
LIB_PATH = os.environ.get("FOVGS_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libfovgs.so")

FOVGS_PS1_OBB = 0
FOVGS_PS1_SUM = 1
FOVGS_PS1_MAX = 2
FOVGS_PS1_LWMC = 3
FOVGS_PS1_VANILLA = 4

FOVGS_VERSION = 202   # include/fovgs.h; every args struct carries it next to its own size (FOVGS_ARGS_HEADER)

_f = C.c_void_p  # all device pointers travel as void*
_HEADER = [("struct_size", C.c_uint32), ("abi_version", C.c_uint32)]


def new_args(cls):
    """An args struct with its two-word header filled in (FOVGS_ARGS_INIT of the C header)."""
    a = cls()
    a.struct_size = C.sizeof(cls)
    a.abi_version = FOVGS_VERSION
    return a



class Camera(C.Structure):
    _fields_ = [
        ("image_height", C.c_int32),
        ("image_width", C.c_int32),
        ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32),
        ("prefiltered", C.c_int32),
        ("debug", C.c_int32),
        ("bg", _f),
        ("viewmatrix", _f),
        ("projmatrix", _f),
        ("campos", _f),
    ]


class FrameStats(C.Structure):
    _fields_ = [
        ("num_rendered", C.c_uint32),
        ("overflow", C.c_uint32),
        ("num_visible", C.c_uint32),
        ("num_blend_tiles", C.c_uint32),
        ("max_tile_instances", C.c_uint32),
        ("reserved", C.c_uint32 * 11),
    ]


class FovFwdArgs(C.Structure):
    _fields_ = _HEADER + [
        ("cam", Camera),
        ("P", C.c_int32),
        ("M_rest", C.c_int32),
        ("means3D", _f),
        ("opacities", _f),
        ("scales", _f),
        ("rotations", _f),
        ("shs_rest", _f),
        ("shs_dcs", _f),
        ("highest_levels", _f),
        ("gaze", _f),
        ("alpha", C.c_float),
        ("blending", C.c_int32),
        ("out_color", _f),
        ("radii", _f),
        ("workspace", _f),
        ("workspace_bytes", C.c_size_t),
        ("max_instances", C.c_int64),
        ("out_point_list", _f),
        ("out_ranges", _f),
        ("packed_color_rows", _f),
        ("early_stats_host", _f),
        ("early_stats_event", _f),
        ("out_color_u8", _f),
    ]


class SmfrFwdArgs(C.Structure):
    _fields_ = _HEADER + [
        ("cam", Camera),
        ("P", C.c_int32),
        ("M", C.c_int32),
        ("means3D", _f),
        ("opacities", _f),
        ("scales", _f),
        ("rotations", _f),
        ("shs", _f),
        ("highest_levels", _f),
        ("gaze", _f),
        ("alpha", C.c_float),
        ("blending", C.c_int32),
        ("out_color", _f),
        ("radii", _f),
        ("workspace", _f),
        ("workspace_bytes", C.c_size_t),
        ("max_instances", C.c_int64),
        ("out_point_list", _f),
        ("out_ranges", _f),
        ("early_stats_host", _f),
        ("early_stats_event", _f),
    ]


class MmfrFwdArgs(C.Structure):
    _fields_ = _HEADER + [
        ("cam", Camera),
        ("P", C.c_int32),
        ("M", C.c_int32),
        ("means3D", _f),
        ("opacities", _f),
        ("scales", _f),
        ("rotations", _f),
        ("shs", _f),
        ("cur_level", C.c_float),
        ("gaze", _f),
        ("alpha", C.c_float),
        ("blending", C.c_int32),
        ("out_color", _f),
        ("radii", _f),
        ("workspace", _f),
        ("workspace_bytes", C.c_size_t),
        ("max_instances", C.c_int64),
        ("out_point_list", _f),
        ("out_ranges", _f),
        ("early_stats_host", _f),
        ("early_stats_event", _f),
    ]


class Ps1FwdArgs(C.Structure):
    _fields_ = _HEADER + [
        ("cam", Camera),
        ("mode", C.c_int32),
        ("P", C.c_int32),
        ("M", C.c_int32),
        ("means3D", _f),
        ("opacities", _f),
        ("scales", _f),
        ("rotations", _f),
        ("cov3D_precomp", _f),
        ("shs", _f),
        ("colors_precomp", _f),
        ("out_color", _f),
        ("radii", _f),
        ("gaussians_count", _f),
        ("contributions", _f),
        ("workspace", _f),
        ("workspace_bytes", C.c_size_t),
        ("max_instances", C.c_int64),
        ("out_point_list", _f),
        ("out_ranges", _f),
        ("loss_map", _f),
        ("early_stats_host", _f),
        ("early_stats_event", _f),
    ]


class Ps1BwdArgs(C.Structure):
    _fields_ = _HEADER + [
        ("cam", Camera),
        ("P", C.c_int32),
        ("M", C.c_int32),
        ("means3D", _f),
        ("scales", _f),
        ("rotations", _f),
        ("cov3D_precomp", _f),
        ("shs", _f),
        ("colors_precomp", _f),
        ("radii", _f),
        ("dL_dout_color", _f),
        ("workspace", _f),
        ("workspace_bytes", C.c_size_t),
        ("max_instances", C.c_int64),
        ("dL_dmeans2D", _f),
        ("dL_dconic", _f),
        ("dL_dopacity", _f),
        ("dL_dcolors", _f),
        ("dL_dmeans3D", _f),
        ("dL_dcov3D", _f),
        ("dL_dsh", _f),
        ("dL_dscales", _f),
        ("dL_drotations", _f),
    ]


class AdamGroup(C.Structure):
    _fields_ = [
        ("param", _f),
        ("grad", _f),
        ("exp_avg", _f),
        ("exp_avg_sq", _f),
        ("n", C.c_int64),
        ("step", C.c_int64),
        ("lr", C.c_double),
        ("beta1", C.c_double),
        ("beta2", C.c_double),
        ("eps", C.c_double),
    ]


ADAM_MAX_GROUPS = 8

# every symbol include/fovgs.h declares (tests/test_abi.py checks this list against the header)
EXPORTS = (
    "fovgs_workspace_bytes",
    "fovgs_forward_fov",
    "fovgs_pack_color_rows",
    "fovgs_forward_smfr",
    "fovgs_forward_mmfr",
    "fovgs_forward_ps1",
    "fovgs_backward_ps1",
    "fovgs_mark_visible",
    "fovgs_knn_workspace_bytes",
    "fovgs_knn_mean_dist2",
    "fovgs_activate_forward",
    "fovgs_activate_backward",
    "fovgs_adam_step",
    "fovgs_read_stats_async",
    "fovgs_fov_tile_tables",
    "fovgs_ps1_geometry",
    "fovgs_fov_geometry",
    "fovgs_set_option",
    "fovgs_profile_enable",
    "fovgs_profile_read",
    "fovgs_profile_count",
    "fovgs_profile_read_frame",
    "fovgs_last_error",
    "fovgs_version",
    "fovgs_struct_size",
    "fovgs_debug_expf_mismatches",
)

# fovgs_struct_id -> the ctypes mirror in this file
STRUCT_IDS = {0: Camera, 1: FrameStats, 2: FovFwdArgs, 3: SmfrFwdArgs, 4: MmfrFwdArgs, 5: Ps1FwdArgs, 6: Ps1BwdArgs, 7: AdamGroup}

_lib = None


def lib():
    """Load libfovgs.so once.  Raises RuntimeError when it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libfovgs.so not found at {LIB_PATH}: build it with `make -C fov-3dgs_b200` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU/PyTorch fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.fovgs_last_error.restype = C.c_char_p
    L.fovgs_version.restype = C.c_int
    L.fovgs_struct_size.restype = C.c_size_t
    L.fovgs_struct_size.argtypes = [C.c_int32]
    if L.fovgs_version() != FOVGS_VERSION:
        raise RuntimeError(f"libfovgs.so at {LIB_PATH} has ABI version {L.fovgs_version()}, this binding was written for {FOVGS_VERSION}: rebuild it")
    for sid, cls in STRUCT_IDS.items():
        if L.fovgs_struct_size(sid) != C.sizeof(cls):
            raise RuntimeError(f"fovgs/_lib.py: {cls.__name__} is {C.sizeof(cls)} bytes, the library's struct is {L.fovgs_struct_size(sid)}")
    L.fovgs_workspace_bytes.restype = C.c_size_t
    L.fovgs_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32]
    L.fovgs_forward_fov.argtypes = [C.POINTER(FovFwdArgs), C.c_void_p]
    L.fovgs_pack_color_rows.argtypes = [C.c_int32, C.c_int32, _f, _f, _f, _f, _f, C.c_void_p]
    L.fovgs_pack_color_rows.restype = C.c_int
    L.fovgs_forward_smfr.argtypes = [C.POINTER(SmfrFwdArgs), C.c_void_p]
    L.fovgs_forward_mmfr.argtypes = [C.POINTER(MmfrFwdArgs), C.c_void_p]
    L.fovgs_forward_ps1.argtypes = [C.POINTER(Ps1FwdArgs), C.c_void_p]
    L.fovgs_backward_ps1.argtypes = [C.POINTER(Ps1BwdArgs), C.c_void_p]
    L.fovgs_mark_visible.argtypes = [C.c_int32, _f, _f, _f, _f, C.c_void_p]
    L.fovgs_knn_workspace_bytes.restype = C.c_size_t
    L.fovgs_knn_workspace_bytes.argtypes = [C.c_int32]
    L.fovgs_knn_mean_dist2.restype = C.c_int
    L.fovgs_knn_mean_dist2.argtypes = [C.c_int32, _f, _f, _f, C.c_size_t, C.c_void_p]
    L.fovgs_activate_forward.argtypes = [C.c_int32, _f, _f, _f, _f, _f, _f, C.c_void_p]
    L.fovgs_activate_forward.restype = C.c_int
    L.fovgs_activate_backward.argtypes = [C.c_int32, _f, _f, _f, _f, _f, _f, _f, _f, _f, C.c_void_p]
    L.fovgs_activate_backward.restype = C.c_int
    L.fovgs_adam_step.argtypes = [C.POINTER(AdamGroup), C.c_int32, C.c_void_p]
    L.fovgs_adam_step.restype = C.c_int
    L.fovgs_read_stats_async.argtypes = [_f, _f, C.c_void_p]
    L.fovgs_fov_tile_tables.argtypes = [_f, C.c_int32, C.c_int32, _f, _f, _f, _f, _f, C.c_void_p]
    L.fovgs_ps1_geometry.argtypes = [_f, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f, _f, _f, _f, _f, C.c_void_p]
    L.fovgs_fov_geometry.argtypes = [_f, C.c_int32, C.c_int32, C.c_int32, _f, _f, _f, _f, C.c_void_p]
    L.fovgs_debug_expf_mismatches.argtypes = [C.c_uint32, C.c_uint32, _f, C.c_void_p]
    L.fovgs_debug_expf_mismatches.restype = C.c_int
    L.fovgs_set_option.argtypes = [C.c_int32, C.c_int32]
    L.fovgs_set_option.restype = C.c_int
    L.fovgs_profile_enable.argtypes = [C.c_int32]
    L.fovgs_profile_enable.restype = C.c_int
    L.fovgs_profile_read.argtypes = [C.POINTER(C.c_float), C.c_int32]
    L.fovgs_profile_read.restype = C.c_int
    L.fovgs_profile_count.restype = C.c_int
    L.fovgs_profile_read_frame.argtypes = [C.c_int32, C.POINTER(C.c_float), C.c_int32]
    L.fovgs_profile_read_frame.restype = C.c_int
    for fn in ("fovgs_forward_fov", "fovgs_forward_smfr", "fovgs_forward_mmfr", "fovgs_forward_ps1", "fovgs_backward_ps1", "fovgs_mark_visible",
               "fovgs_read_stats_async", "fovgs_fov_tile_tables", "fovgs_ps1_geometry", "fovgs_fov_geometry"):
        getattr(L, fn).restype = C.c_int
    _lib = L
    return L


def check(status, what):
    if status != 0:
        msg = lib().fovgs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")
