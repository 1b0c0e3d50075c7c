"""The binding a reference maintainer would add (INTEGRATION.md section 2 quotes this file verbatim between its
`# --- binding stub` markers; tests/test_abi.py checks that, and checks every ctypes mirror below against the library's
fovgs_struct_size()).  Replaces the `_C.rasterize_gaussians(*args)` call of
FOV/diff_gaussian_rasterization_fov_pcheck_obb/__init__.py:120.  FOVGS_LIB = path of libfovgs.so (default: next to this repo's package)."""
import os
# --- binding stub (begin)
import ctypes as C, torch
lib = C.CDLL(os.environ.get("FOVGS_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fov-3dgs_b200", "lib", "libfovgs.so")))

class fovgs_camera(C.Structure):                       # include/fovgs.h: fovgs_camera
    _fields_ = [("image_height", C.c_int32), ("image_width", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("scale_modifier", C.c_float), ("sh_degree", C.c_int32), ("prefiltered", C.c_int32), ("debug", C.c_int32),
                ("bg", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p)]

class fovgs_fov_fwd_args(C.Structure):                 # include/fovgs.h: fovgs_fov_fwd_args
    _fields_ = [("struct_size", C.c_uint32), ("abi_version", C.c_uint32),          # FOVGS_ARGS_HEADER
                ("cam", fovgs_camera), ("P", C.c_int32), ("M_rest", C.c_int32),
                ("means3D", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
                ("shs_rest", C.c_void_p), ("shs_dcs", C.c_void_p), ("highest_levels", C.c_void_p), ("gaze", C.c_void_p),
                ("alpha", C.c_float), ("blending", C.c_int32), ("out_color", C.c_void_p), ("radii", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("max_instances", C.c_int64),
                ("out_point_list", C.c_void_p), ("out_ranges", C.c_void_p), ("packed_color_rows", C.c_void_p),
                ("early_stats_host", C.c_void_p), ("early_stats_event", C.c_void_p), ("out_color_u8", C.c_void_p)]

lib.fovgs_workspace_bytes.restype = C.c_size_t
lib.fovgs_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32]
lib.fovgs_last_error.restype = C.c_char_p
lib.fovgs_struct_size.restype = C.c_size_t
FOVGS_VERSION = 202                                     # include/fovgs.h
assert lib.fovgs_version() == FOVGS_VERSION
assert lib.fovgs_struct_size(0) == C.sizeof(fovgs_camera) and lib.fovgs_struct_size(2) == C.sizeof(fovgs_fov_fwd_args)

def rasterize_gaussians_fov(shs_dcs, highest_levels, gazeArray, alpha, blending, bg, means3D, colors, opacity, scales, rotations,
                            scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, H, W, shs_rest, degree,
                            campos, prefiltered, debug, cap=None):
    """Same 24 arguments as FOV/rasterize_points.h:17-44; returns (num_rendered, color, radii)."""
    P = means3D.shape[0]
    cap = cap or max(1 << 20, 8 * P)
    nbytes = lib.fovgs_workspace_bytes(P, W, H, cap, 1, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=means3D.device)        # torch's allocator owns all memory
    color = torch.empty((3, H, W), dtype=torch.float32, device=means3D.device)
    radii = torch.empty((P,), dtype=torch.int32, device=means3D.device)
    a = fovgs_fov_fwd_args()
    a.struct_size, a.abi_version = C.sizeof(fovgs_fov_fwd_args), FOVGS_VERSION
    a.cam = fovgs_camera(H, W, tanfovx, tanfovy, scale_modifier, degree, int(prefiltered), int(debug),
                         bg.data_ptr(), viewmatrix.data_ptr(), projmatrix.data_ptr(), campos.data_ptr())
    a.P, a.M_rest = P, (shs_rest.shape[1] if shs_rest.numel() else 0)
    for name, t in (("means3D", means3D), ("opacities", opacity), ("scales", scales), ("rotations", rotations),
                    ("shs_rest", shs_rest), ("shs_dcs", shs_dcs), ("highest_levels", highest_levels), ("gaze", gazeArray)):
        setattr(a, name, t.contiguous().data_ptr() if t.numel() else None)
    a.alpha, a.blending = alpha, int(bool(blending))
    a.out_color, a.radii = color.data_ptr(), radii.data_ptr()
    a.workspace, a.workspace_bytes, a.max_instances = ws.data_ptr(), nbytes, cap
    st = lib.fovgs_forward_fov(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if st != 0:
        raise RuntimeError(lib.fovgs_last_error().decode())
    stats = torch.empty(16, dtype=torch.int32).pin_memory()                   # fovgs_frame_stats
    lib.fovgs_read_stats_async(C.c_void_p(ws.data_ptr()), C.c_void_p(stats.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.current_stream().synchronize()
    if int(stats[1]):                                                          # overflow: re-run with the exact size
        return rasterize_gaussians_fov(shs_dcs, highest_levels, gazeArray, alpha, blending, bg, means3D, colors, opacity, scales,
                                       rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
                                       shs_rest, degree, campos, prefiltered, debug, cap=int(stats[0]) * 5 // 4 + 1024)
    return int(stats[0]) & 0xFFFFFFFF, color, radii


# ---- optimizer hook: fovgs_adam_step (scene/gaussian_model.py:289 / eff_finetune.py:146)
class fovgs_adam_group(C.Structure):                   # include/fovgs.h: fovgs_adam_group
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("step", C.c_int64), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double)]

def adam_step(optimizer):                              # one launch for all (<= 8) parameter groups
    gs = []
    for group in optimizer.param_groups:
        p = group["params"][0]
        if p.grad is None: continue
        st = optimizer.state[p]
        if not st:
            st["step"], st["exp_avg"], st["exp_avg_sq"] = torch.tensor(0.0), torch.zeros_like(p), torch.zeros_like(p)
        st["step"] += 1
        gs.append(fovgs_adam_group(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                   p.numel(), int(st["step"]), group["lr"], *group["betas"], group["eps"]))
    arr = (fovgs_adam_group * len(gs))(*gs)
    if lib.fovgs_adam_step(arr, len(gs), C.c_void_p(torch.cuda.current_stream().cuda_stream)) != 0:
        raise RuntimeError(lib.fovgs_last_error().decode())
# --- binding stub (end)
