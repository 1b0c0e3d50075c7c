"""ctypes wrapper of the CPU oracle (oracle/fovgs_oracle.c).   *** TEST INFRASTRUCTURE ONLY ***

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).  The product
package fov-3dgs_b200/ never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class Camera(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("scale_modifier", C.c_float), ("sh_degree", C.c_int32), ("bg", C.c_float * 3),
                ("view", C.c_float * 16), ("proj", C.c_float * 16), ("campos", C.c_float * 3)]


def build(force=False):
    src = os.path.join(HERE, "fovgs_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B" if force else "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_forward_ps1.restype = C.c_int64
        _lib.orc_forward_fov.restype = C.c_int64
        _lib.orc_forward_smfr.restype = C.c_int64
        _lib.orc_forward_mmfr.restype = C.c_int64
        _lib.orc_backward_ps1.restype = C.c_int
        _lib.orc_ambiguous.restype = C.c_int64
    return _lib


def set_ambiguity(on):
    """Binning passes record the (tile, Gaussian) OBB decisions that depend on the last place of the eigenvector
    normalisation (GPU: MUFU.RSQ, here: 1/sqrtf) — see the header of fovgs_oracle.c.  Read them with `ambiguous()`."""
    lib().orc_set_ambiguity(1 if on else 0)


def ambiguous():
    """int64 keys (tile << 32 | Gaussian id) of the last binning pass's rsqrt-sensitive decisions (needs set_ambiguity(True))."""
    L = lib()
    n = int(L.orc_ambiguous(None, 0))
    out = np.zeros(max(n, 1), np.uint64)
    L.orc_ambiguous(out.ctypes.data_as(C.c_void_p), C.c_int64(n))
    return out[:n].astype(np.int64)


def instance_keys(point_list, ranges):
    """int64 keys (tile << 32 | Gaussian id) of a sorted point list with its per-tile ranges."""
    rg = np.asarray(ranges, np.int64)
    n = np.maximum(rg[:, 1] - rg[:, 0], 0)
    tile = np.repeat(np.arange(rg.shape[0], dtype=np.int64), n)
    return (tile << 32) | np.asarray(point_list, np.int64)[: tile.size]


def _cam(cam, sh_degree, bg=(0.0, 0.0, 0.0), scale_modifier=1.0):
    c = Camera()
    c.W = int(cam["image_width"]); c.H = int(cam["image_height"])
    c.tanfovx = float(cam["tanfovx"]); c.tanfovy = float(cam["tanfovy"])
    c.scale_modifier = float(scale_modifier); c.sh_degree = int(sh_degree)
    c.bg[:] = [float(b) for b in bg]
    c.view[:] = np.asarray(cam["viewmatrix"], np.float32).ravel().tolist()
    c.proj[:] = np.asarray(cam["projmatrix"], np.float32).ravel().tolist()
    c.campos[:] = np.asarray(cam["campos"], np.float32).ravel().tolist()
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


PS1_MODES = {"obb": 0, "sum": 1, "max": 2, "lwmc": 3, "vanilla": 4}


def forward_ps1(scene, cam, mode="obb", bg=(0.0, 0.0, 0.0), list_cap=None, loss_map=None):
    """mode: 'obb' | 'sum' | 'max' | 'lwmc' (loss_map [H,W] required) | 'vanilla' (the stock diff-gaussian-rasterization:
    no OBB test, no -4.5 cut, no statistics).  Returns dict with color, radii, num_rendered,
    point_list, ranges, means2D, depths, conic, cov3D, rgb, clamped (+ gaussians_count, contributions, final_T,
    n_contrib for the training-family modes)."""
    L = lib()
    P = scene["means3D"].shape[0]
    M = scene["shs"].shape[1]
    W, H = cam["image_width"], cam["image_height"]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    c = _cam(cam, scene["sh_degree"], bg)
    cap = int(list_cap or max(1 << 20, 64 * P))
    o = {
        "color": np.zeros((3, H, W), np.float32), "radii": np.zeros(P, np.int32),
        "gaussians_count": np.zeros(P, np.int32), "contributions": np.zeros(P, np.float32),
        "means2D": np.zeros((P, 2), np.float32), "depths": np.zeros(P, np.float32), "conic": np.zeros((P, 3), np.float32),
        "cov3D": np.zeros((P, 6), np.float32), "rgb": np.zeros((P, 3), np.float32), "clamped": np.zeros((P, 3), np.uint8),
        "point_list": np.zeros(cap, np.uint32), "ranges": np.zeros((T, 2), np.uint32),
        "final_T": np.zeros(H * W, np.float32), "n_contrib": np.zeros(H * W, np.uint32),
    }
    ins = [_f32(scene["means3D"]), _f32(scene["opacity"]), _f32(scene["scales"]), _f32(scene["rotations"]), _f32(scene["shs"])]
    if mode == "lwmc" and loss_map is None:
        raise ValueError("mode 'lwmc' needs loss_map")
    lm = None if loss_map is None else _f32(np.asarray(loss_map).reshape(-1))
    if lm is not None and lm.size != H * W:
        raise ValueError("loss_map must have H*W elements")
    n = L.orc_forward_ps1(C.byref(c), PS1_MODES[mode], P, M, *[_p(a) for a in ins], _p(o["color"]), _p(o["radii"]),
                          _p(o["gaussians_count"]), _p(o["contributions"]), _p(o["means2D"]), _p(o["depths"]), _p(o["conic"]),
                          _p(o["cov3D"]), _p(o["rgb"]), _p(o["clamped"]), _p(o["point_list"]), C.c_int64(cap), _p(o["ranges"]),
                          _p(o["final_T"]), _p(o["n_contrib"]), _p(lm))
    o["num_rendered"] = int(n)
    o["point_list"] = o["point_list"][: min(n, cap)]
    return o


def forward_fov(scene, cam, gaze, alpha=0.05, bg=(0.0, 0.0, 0.0), list_cap=None):
    """scene must carry the foveated tensors (synth.add_foveation)."""
    L = lib()
    P = scene["means3D"].shape[0]
    M_rest = scene["shs_rest"].shape[1]
    W, H = cam["image_width"], cam["image_height"]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    c = _cam(cam, scene["sh_degree"], bg)
    cap = int(list_cap or max(1 << 20, 64 * P))
    o = {
        "color": np.zeros((3, H, W), np.float32), "radii": np.zeros(P, np.int32),
        "means2D": np.zeros((P, 2), np.float32), "depths": np.zeros(P, np.float32), "conic": np.zeros((P, 3), np.float32),
        "point_list": np.zeros(cap, np.uint32), "ranges": np.zeros((T, 2), np.uint32),
        "tile_level": np.zeros(T, np.float32), "tile_min": np.zeros(T, np.float32), "tile_blend": np.zeros(T, np.uint8),
        "level_ranges": np.zeros((P, 2), np.int32),
    }
    g = _f32(np.asarray(gaze, np.float32))
    ins = [_f32(scene["means3D"]), _f32(scene["opacities4"]), _f32(scene["scales"]), _f32(scene["rotations"]),
           _f32(scene["shs_rest"]), _f32(scene["shs_dcs"]), _f32(scene["highest_levels"])]
    n = L.orc_forward_fov(C.byref(c), P, M_rest, *[_p(a) for a in ins], _p(g), C.c_float(alpha), _p(o["color"]), _p(o["radii"]),
                          _p(o["means2D"]), _p(o["depths"]), _p(o["conic"]), _p(o["point_list"]), C.c_int64(cap), _p(o["ranges"]),
                          _p(o["tile_level"]), _p(o["tile_min"]), _p(o["tile_blend"]), _p(o["level_ranges"]))
    o["num_rendered"] = int(n)
    o["point_list"] = o["point_list"][: min(n, cap)]
    return o


def forward_smfr(scene, cam, gaze, alpha=0.05, bg=(0.0, 0.0, 0.0), list_cap=None):
    """SMFR baseline (naive_pcheck_obb): scene carries shs [P,M,3], opacity [P,1] and highest_levels [P,1]."""
    L = lib()
    P = scene["means3D"].shape[0]
    M = scene["shs"].shape[1]
    W, H = cam["image_width"], cam["image_height"]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    c = _cam(cam, scene["sh_degree"], bg)
    cap = int(list_cap or max(1 << 20, 64 * P))
    o = {"color": np.zeros((3, H, W), np.float32), "radii": np.zeros(P, np.int32),
         "point_list": np.zeros(cap, np.uint32), "ranges": np.zeros((T, 2), np.uint32)}
    g = _f32(np.asarray(gaze, np.float32))
    ins = [_f32(scene["means3D"]), _f32(scene["opacity"]), _f32(scene["scales"]), _f32(scene["rotations"]), _f32(scene["shs"]),
           _f32(scene["highest_levels"])]
    n = L.orc_forward_smfr(C.byref(c), P, M, *[_p(a) for a in ins], _p(g), C.c_float(alpha), _p(o["color"]), _p(o["radii"]),
                           _p(o["point_list"]), C.c_int64(cap), _p(o["ranges"]))
    o["num_rendered"] = int(n)
    o["point_list"] = o["point_list"][: min(n, cap)]
    return o


def forward_mmfr(scene, cam, cur_level, gaze, alpha=0.05, bg=(0.0, 0.0, 0.0), list_cap=None):
    """MMFR baseline (mmfr_pcheck_obb), ONE level call: scene carries that level's model (shs [P,M,3], opacity [P,1])."""
    L = lib()
    P = scene["means3D"].shape[0]
    M = scene["shs"].shape[1]
    W, H = cam["image_width"], cam["image_height"]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    c = _cam(cam, scene["sh_degree"], bg)
    cap = int(list_cap or max(1 << 20, 64 * P))
    o = {"color": np.zeros((3, H, W), np.float32), "radii": np.zeros(P, np.int32), "tile_skip": np.zeros(T, np.uint8),
         "point_list": np.zeros(cap, np.uint32), "ranges": np.zeros((T, 2), np.uint32)}
    g = _f32(np.asarray(gaze, np.float32))
    ins = [_f32(scene["means3D"]), _f32(scene["opacity"]), _f32(scene["scales"]), _f32(scene["rotations"]), _f32(scene["shs"])]
    n = L.orc_forward_mmfr(C.byref(c), P, M, *[_p(a) for a in ins], C.c_float(float(cur_level)), _p(g), C.c_float(alpha),
                           _p(o["color"]), _p(o["radii"]), _p(o["point_list"]), C.c_int64(cap), _p(o["ranges"]), _p(o["tile_skip"]))
    o["num_rendered"] = int(n)
    o["point_list"] = o["point_list"][: min(n, cap)]
    return o


def backward_ps1(scene, cam, fwd, dL_dpix, bg=(0.0, 0.0, 0.0), vanilla=False):
    """fwd: result of forward_ps1(mode='sum') — or mode='vanilla' with vanilla=True.  Returns dict of the 8 gradients
    (+ dL_dconic scratch)."""
    L = lib()
    P = scene["means3D"].shape[0]
    M = scene["shs"].shape[1]
    c = _cam(cam, scene["sh_degree"], bg)
    g = {
        "dL_dmeans2D": np.zeros((P, 3), np.float32), "dL_dconic": np.zeros((P, 4), np.float32),
        "dL_dopacity": np.zeros((P, 1), np.float32), "dL_dcolors": np.zeros((P, 3), np.float32),
        "dL_dmeans3D": np.zeros((P, 3), np.float32), "dL_dcov3D": np.zeros((P, 6), np.float32),
        "dL_dsh": np.zeros((P, M, 3), np.float32), "dL_dscales": np.zeros((P, 3), np.float32),
        "dL_drotations": np.zeros((P, 4), np.float32),
    }
    ins = [_f32(scene["means3D"]), _f32(scene["scales"]), _f32(scene["rotations"]), _f32(scene["shs"]), _f32(scene["opacity"])]
    pl = np.ascontiguousarray(fwd["point_list"], np.uint32)
    (L.orc_backward_vanilla if vanilla else L.orc_backward_ps1)(C.byref(c), P, M, *[_p(a) for a in ins], _p(np.ascontiguousarray(fwd["radii"], np.int32)),
                       _p(_f32(fwd["means2D"])), _p(_f32(fwd["conic"])), _p(_f32(fwd["rgb"])),
                       _p(np.ascontiguousarray(fwd["clamped"], np.uint8)), _p(_f32(fwd["cov3D"])), _p(pl),
                       _p(np.ascontiguousarray(fwd["ranges"], np.uint32)), _p(_f32(fwd["final_T"])),
                       _p(np.ascontiguousarray(fwd["n_contrib"], np.uint32)), _p(_f32(dL_dpix)),
                       _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]),
                       _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def tile_tables(W, H, gaze, alpha=0.05):
    L = lib()
    T = ((W + 15) // 16) * ((H + 15) // 16)
    lvl, mn, gx, gy = (np.zeros(T, np.float32) for _ in range(4))
    bl = np.zeros(T, np.uint8)
    g = _f32(np.asarray(gaze, np.float32))
    L.orc_tile_tables(int(W), int(H), _p(g), C.c_float(alpha), _p(lvl), _p(mn), _p(gx), _p(gy), _p(bl))
    return {"tile_level": lvl, "tile_min": mn, "grad_x": gx, "grad_y": gy, "blending": bl}
