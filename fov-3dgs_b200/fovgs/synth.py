"""Seeded synthetic scenes and cameras (SURVEY.md §8d / BASELINE.md §3): there is no dataset or checkpoint in
this environment, so tests, goldens and bench.py render Gaussians of the reference's tensor layout drawn from the
distributions written down in BASELINE.md.  Camera matrices follow the reference's own formulas
(utils/graphics_utils.py:38-71 getWorld2View2 / getProjectionMatrix, scene/cameras.py:54-57: matrices are stored
TRANSPOSED, full_proj = world_view^T-convention product, camera_center = inverse(world_view)[3,:3]).

Everything is numpy (float64 math, float32 results) so the same bytes are produced on every machine.
"""
import math

import numpy as np

LEVEL_FRACTIONS = (0.599, 0.183, 0.043, 0.175)  # share of Gaussians whose highest level is 0,1,2,3 (pnum/ours-Q/bicycle.txt)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _finish(rng, xyz, log_scale_mu, log_scale_sigma, op_mu, op_sigma, sh_degree=3, anisotropic=True):
    P = xyz.shape[0]
    ls = rng.normal(log_scale_mu, log_scale_sigma, size=(P, 1)) + rng.normal(0.0, 0.25, size=(P, 3))
    if anisotropic:
        ax = rng.integers(0, 3, size=P)
        ls[np.arange(P), ax] += math.log(0.3)
    scales = np.exp(ls)
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opacity = _sigmoid(rng.normal(op_mu, op_sigma, size=(P, 1)))
    M = (sh_degree + 1) ** 2
    shs = np.concatenate([rng.normal(0.0, 1.0, size=(P, 1, 3)), rng.normal(0.0, 0.1, size=(P, M - 1, 3))], axis=1)
    return {
        "means3D": xyz.astype(np.float32),
        "scales": scales.astype(np.float32),
        "rotations": q.astype(np.float32),
        "opacity": opacity.astype(np.float32),
        "shs": shs.astype(np.float32),
        "sh_degree": sh_degree,
    }


def make_scene_cube(P=10000, seed=0):
    """Config 1: xyz ~ U([-1,1]^3), log_scale ~ N(-3.0, 0.5), opacity = sigmoid(N(0,1.5))."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1.0, 1.0, size=(P, 3))
    return _finish(rng, xyz, -3.0, 0.5, 0.0, 1.5, anisotropic=False)


def make_scene_bicycle(P=6_000_000, seed=1, log_scale_mu=-4.6):
    """Configs 2-5: 55 % ground disk, 25 % central blob, 20 % background shell (y is up)."""
    rng = np.random.default_rng(seed)
    n_g = int(P * 0.55)
    n_c = int(P * 0.25)
    n_b = P - n_g - n_c
    r = 8.0 * np.sqrt(rng.uniform(0, 1, n_g))
    th = rng.uniform(0, 2 * math.pi, n_g)
    ground = np.stack([r * np.cos(th), 0.05 * r * rng.normal(0, 1, n_g) * 0.5, r * np.sin(th)], axis=1)
    blob = rng.normal(0.0, 0.6, size=(n_c, 3))
    blob[:, 1] = np.abs(blob[:, 1]) * 0.8
    rb = rng.uniform(8.0, 30.0, n_b)
    d = rng.normal(size=(n_b, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:, 1] = np.abs(d[:, 1]) * 0.6
    shell = d * rb[:, None]
    xyz = np.concatenate([ground, blob, shell], axis=0)
    perm = rng.permutation(P)  # real scenes are not spatially sorted
    xyz = xyz[perm]
    scene = _finish(rng, xyz, log_scale_mu, 0.9, 0.5, 2.0)
    # shell Gaussians are larger (distant content is coarser), keeps the far field covered
    far = np.linalg.norm(xyz, axis=1) > 8.0
    scene["scales"][far] *= 4.0
    return scene


def add_foveation(scene, seed=7):
    """Foveated inputs (Appendix B of SURVEY.md / compose_models.py:39-80): highest_levels [P,1] float,
    opacities [P,4] and shs_dcs [P,4,3] with the carry rule (level i keeps level i-1's value unless re-drawn)."""
    rng = np.random.default_rng(seed)
    P = scene["means3D"].shape[0]
    u = rng.uniform(0, 1, P)
    edges = np.cumsum(LEVEL_FRACTIONS)
    hl = np.searchsorted(edges, u, side="right").clip(0, 3).astype(np.float32)
    op = np.repeat(scene["opacity"].astype(np.float64), 4, axis=1)
    dc = np.repeat(scene["shs"][:, 0:1, :].astype(np.float64), 4, axis=1)
    for lvl in range(1, 4):
        member = hl >= lvl
        redraw = member & (rng.uniform(0, 1, P) < 0.3)
        op[:, lvl] = op[:, lvl - 1]
        dc[:, lvl] = dc[:, lvl - 1]
        k = int(redraw.sum())
        op[redraw, lvl] = np.clip(op[redraw, lvl - 1] * rng.uniform(0.8, 1.6, k), 0.0, 1.0)
        dc[redraw, lvl] = dc[redraw, lvl - 1] + rng.normal(0, 0.2, size=(k, 3))
    out = dict(scene)
    out["highest_levels"] = hl.reshape(P, 1)
    out["opacities4"] = op.astype(np.float32)
    out["shs_dcs"] = dc.astype(np.float32)
    out["shs_rest"] = np.ascontiguousarray(scene["shs"][:, 1:, :])
    return out


def look_at_camera(W, H, fovx_deg, eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), znear=0.01, zfar=100.0):
    """Camera dict with the reference's matrices.  View space: +z forward, +y down (COLMAP convention)."""
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    up = np.asarray(up, dtype=np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, up)  # with y-down image axes: right = fwd x up_world  (x to the right)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], axis=1)  # camera-to-world rotation (columns = camera axes)
    t = -R.T @ eye
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.T
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    fovx = math.radians(fovx_deg)
    fovy = 2.0 * math.atan(math.tan(fovx / 2.0) * H / W)
    tanx, tany = math.tan(fovx / 2.0), math.tan(fovy / 2.0)
    top, right_ = tany * znear, tanx * znear
    Pm = np.zeros((4, 4))
    Pm[0, 0] = 2.0 * znear / (2 * right_)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    wv = np.float32(Rt).T.copy()                  # world_view_transform (transposed)
    pj = np.float32(Pm).T.copy()                  # projection_matrix (transposed)
    full = (wv.astype(np.float32) @ pj.astype(np.float32)).astype(np.float32)
    campos = np.linalg.inv(wv.astype(np.float64))[3, :3].astype(np.float32)
    return {
        "image_width": int(W), "image_height": int(H),
        "tanfovx": float(tanx), "tanfovy": float(tany),
        "viewmatrix": np.ascontiguousarray(wv), "projmatrix": np.ascontiguousarray(full),
        "campos": np.ascontiguousarray(campos), "FoVx": fovx, "FoVy": fovy,
    }


def ring_cameras(n=30, W=1920, H=1080, fovx_deg=60.9, radius=4.5, height=1.2):
    cams = []
    for i in range(n):
        a = 2 * math.pi * i / n
        cams.append(look_at_camera(W, H, fovx_deg, (radius * math.cos(a), height, radius * math.sin(a)), (0.0, 0.3, 0.0)))
    return cams


def config1_camera():
    """256x256, FoV 60 deg, camera at z = -4 looking at the origin."""
    return look_at_camera(256, 256, 60.0, (0.0, 0.0, -4.0))


GAZES_9 = [(0.25 * i, 0.25 * j) for i in range(1, 4) for j in range(1, 4)]  # render_compose_gazes_fps.py:26
