"""GPU parity tests (pytest -m gpu): the sm_100a library, called through the C-ABI by the drop-in Python surface,
against (1) golden vectors of the unmodified reference CUDA, (2) the CPU oracle on fresh seeded inputs,
(3) size-independent properties at BASELINE.json's full size (6 M Gaussians, 1920x1080).

Bars: integer / index outputs bit-exact; image within 1e-4 absolute per pixel (measured 0.0 vs the
reference binary); gradients rel-L2 <= 1e-4.
"""
import os

import numpy as np
import pytest
import torch

from fovgs import ops, synth

pytestmark = pytest.mark.gpu
IMG_TOL = 1e-4


def _cuda(d):
    return {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def _settings(mod, cam, sh_degree, bg=None, debug=False):
    c = _cuda(cam)
    bg = torch.zeros(3, device="cuda") if bg is None else bg
    return mod.GaussianRasterizationSettings(
        image_height=cam["image_height"], image_width=cam["image_width"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
        bg=bg, scale_modifier=1.0, viewmatrix=c["viewmatrix"], projmatrix=c["projmatrix"], sh_degree=sh_degree,
        campos=c["campos"], prefiltered=False, debug=debug)


def _g(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        pytest.skip(f"golden fixture {name} missing")
    return np.load(p)


def _run_ps1(mode, scene, cam, want_lists=True):
    import diff_gaussian_rasterization_pcheck_obb as m
    sc = _cuda(scene)
    rs = _settings(m, cam, scene["sh_degree"])
    return ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs,
                           want_lists=want_lists), sc, rs


def _run_fov(scene_f, cam, gaze, want_lists=True):
    import diff_gaussian_rasterization_fov_pcheck_obb as m
    sc = _cuda(scene_f)
    rs = _settings(m, cam, scene_f["sh_degree"])
    g = torch.tensor(np.asarray(gaze, np.float32)).cuda()
    return ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"],
                           sc["highest_levels"], g, 0.05, True, rs, want_lists=want_lists), sc, rs


# ---------------------------------------------------------------------------------------------------------------
# 1. golden vectors of the reference CUDA (config 1)
# ---------------------------------------------------------------------------------------------------------------
def test_obb_matches_reference_golden(scene_small, golden_dir):
    g = _g(golden_dir, "obb_small_c0.npz")
    s, c = scene_small
    (n, color, radii, item, pl, rg), _, _ = _run_ps1(ops.MODE_OBB, s, c)
    assert n == int(g["num_rendered"])
    assert np.array_equal(radii.cpu().numpy(), g["radii"])
    assert np.array_equal(pl.cpu().numpy(), g["point_list"])
    assert np.array_equal(rg.cpu().numpy(), g["ranges"])
    geo = ops.geometry(item, ops.MODE_OBB, len(g["radii"]), c["image_width"], c["image_height"])
    vis = g["radii"] > 0
    for k in ("means2D", "depths", "conic"):
        a = geo[k].cpu().numpy().view(np.int32).reshape(len(vis), -1)[vis]
        b = np.ascontiguousarray(g[k]).view(np.int32).reshape(len(vis), -1)[vis]
        assert np.array_equal(a, b), k
    assert np.abs(color.cpu().numpy() - g["color"]).max() <= IMG_TOL


def test_sum_forward_and_backward_match_reference_golden(scene_small, golden_dir):
    g = _g(golden_dir, "sum_small_c0.npz")
    gb = _g(golden_dir, "sum_small_c0_bwd.npz")
    s, c = scene_small
    (n, color, radii, item, gcount, contrib, pl, rg), sc, rs = _run_ps1(ops.MODE_SUM, s, c)
    assert n == int(g["num_rendered"])
    assert np.array_equal(pl.cpu().numpy(), g["point_list"])
    assert np.array_equal(gcount.cpu().numpy(), g["gaussians_count"])
    rel = np.abs(contrib.cpu().numpy() - g["contributions"]) / (np.abs(g["contributions"]) + 1e-6)
    assert rel.max() <= 1e-3
    assert np.abs(color.cpu().numpy() - g["color"]).max() <= IMG_TOL
    if "numpy_grad" not in gb.files:
        pytest.skip("backward golden predates numpy-seeded dL/dpixel")
    H, W = c["image_height"], c["image_width"]
    grad_out = torch.from_numpy(np.random.default_rng(int(gb["grad_seed"])).standard_normal((3, H, W)).astype(np.float32)).cuda()
    grads = ops.backward_ps1(item, sc["means3D"], radii, sc["scales"], sc["rotations"], None, sc["shs"], None, rs, grad_out)
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for nm, t in zip(names, grads):
        a, b = t.cpu().numpy().ravel(), gb[nm].ravel()
        assert np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30) <= 1e-4, nm


def test_vanilla_forward_and_backward_match_reference_golden(scene_small, golden_dir):
    """diff_gaussian_rasterization (the stock rasterizer, cuda_type="original"): whole-rectangle lists, no -4.5 cut — against
    the unmodified reference binary (oracle/_ref/ref_vanilla_C), full-sort path and lazy path."""
    g = _g(golden_dir, "vanilla_small_c0.npz")
    gb = _g(golden_dir, "vanilla_small_c0_bwd.npz")
    s, c = scene_small
    (n, color, radii, item, pl, rg), sc, rs = _run_ps1(ops.MODE_VANILLA, s, c)
    assert n == int(g["num_rendered"])
    assert np.array_equal(radii.cpu().numpy(), g["radii"])
    assert np.array_equal(pl.cpu().numpy(), g["point_list"])
    assert np.array_equal(rg.cpu().numpy(), g["ranges"])
    assert np.abs(color.cpu().numpy() - g["color"]).max() <= IMG_TOL
    (n2, color2, radii2, item2), _, _ = _run_ps1(ops.MODE_VANILLA, s, c, want_lists=False)      # lazy sort + blend
    assert n2 == n and torch.equal(color2, color) and torch.equal(radii2, radii)
    H, W = c["image_height"], c["image_width"]
    grad_out = torch.from_numpy(np.random.default_rng(int(gb["grad_seed"])).standard_normal((3, H, W)).astype(np.float32)).cuda()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for it, rd in ((item, radii), (item2, radii2)):
        grads = ops.backward_ps1(it, sc["means3D"], rd, sc["scales"], sc["rotations"], None, sc["shs"], None, rs, grad_out)
        for nm, t in zip(names, grads):
            a, b = t.cpu().numpy().ravel(), gb[nm].ravel()
            assert np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30) <= 1e-4, nm


def test_vanilla_vs_oracle_and_drop_in_package():
    """Fresh inputs against the CPU oracle (lists bit-exact, image within tolerance, gradients rel-L2), then the package the
    reference imports (fov3dgs/gaussian_wrapper.py:2): returns (color, radii), differentiable."""
    import oracle
    import diff_gaussian_rasterization as mv
    s = synth.make_scene_cube(3000, 41)
    c = _small_cam(176, 112)
    o = oracle.forward_ps1(s, c, "vanilla")
    (n, color, radii, item, pl, rg), sc, rs = _run_ps1(ops.MODE_VANILLA, s, c)
    assert n == o["num_rendered"] and n > oracle.forward_ps1(s, c, "sum")["num_rendered"]
    assert np.array_equal(radii.cpu().numpy(), o["radii"])
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
    assert np.array_equal(rg.cpu().numpy().astype(np.uint32), o["ranges"])
    assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL
    grad_out = np.random.default_rng(2).standard_normal((3, 112, 176)).astype(np.float32)
    go = oracle.backward_ps1(s, c, o, grad_out, vanilla=True)
    g = ops.backward_ps1(item, sc["means3D"], radii, sc["scales"], sc["rotations"], None, sc["shs"], None, rs,
                         torch.from_numpy(grad_out).cuda())
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for nm, t in zip(names, g):
        a, b = t.cpu().numpy().ravel(), go[nm].ravel()
        assert np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30) <= 1e-4, nm
    r = mv.GaussianRasterizer(raster_settings=_settings(mv, c, 3))
    means = sc["means3D"].clone().requires_grad_(True)
    out = r(means3D=means, means2D=torch.zeros_like(means), opacities=sc["opacity"], shs=sc["shs"], scales=sc["scales"],
            rotations=sc["rotations"])
    assert len(out) == 2 and torch.equal(out[0], color) and torch.equal(out[1], radii)
    (out[0] * torch.from_numpy(grad_out).cuda()).sum().backward()
    assert np.linalg.norm(means.grad.cpu().numpy().ravel() - go["dL_dmeans3D"].ravel()) <= 1e-4 * np.linalg.norm(go["dL_dmeans3D"].ravel())
    # a SUM frame on the same pooled workspace afterwards must get its -4.5 cut back (the cut travels in the frame header)
    os_ = oracle.forward_ps1(s, c, "sum")
    (ns, cs, *_), _, _ = _run_ps1(ops.MODE_SUM, s, c)
    assert ns == os_["num_rendered"] and np.abs(cs.cpu().numpy() - os_["color"]).max() <= IMG_TOL


@pytest.mark.parametrize("variant", ["max", "lwmc"])
def test_pruning_variants_match_reference_golden(scene_small, golden_dir, variant):
    g = _g(golden_dir, f"{variant}_small_c0.npz")
    s, c = scene_small
    sc = _cuda(s)
    import diff_gaussian_rasterization_pcheck_obb_sum as m
    rs = _settings(m, c, s["sh_degree"])
    H, W = c["image_height"], c["image_width"]
    lm = None
    if variant == "lwmc":
        lm = torch.from_numpy(np.random.default_rng(int(g["loss_map_seed"])).random((H, W)).astype(np.float32)).cuda()
    mode = ops.MODE_MAX if variant == "max" else ops.MODE_LWMC
    for lazy in (False, True):
        r = ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs,
                            want_lists=not lazy, loss_map=lm)
        n, color, radii, item, gcount, contrib = r[:6]
        assert n == int(g["num_rendered"])
        assert np.array_equal(gcount.cpu().numpy(), g["gaussians_count"])
        assert np.array_equal(color.cpu().numpy(), g["color"])          # measured: bit-identical to the reference binary
        if variant == "max":
            assert np.array_equal(contrib.cpu().numpy(), g["contributions"])   # exact maximum, exact counts
        else:
            a, b = contrib.cpu().numpy().astype(np.float64), g["contributions"].astype(np.float64)
            assert (np.abs(a - b) / (np.abs(b) + 1e-3)).max() <= 1e-4


@pytest.mark.parametrize("gi", [0, 1])
def test_smfr_matches_reference_golden(scene_small, golden_dir, gi):
    g = _g(golden_dir, f"smfr_small_c0_g{gi}.npz")
    s, c = scene_small
    sc = _cuda(synth.add_foveation(s))
    import diff_gaussian_rasterization_naive_pcheck_obb as m
    rs = _settings(m, c, s["sh_degree"])
    gz = torch.from_numpy(np.asarray(g["gaze"], np.float32)).cuda()
    n, color, radii, pl, rg, item = ops.forward_smfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                     sc["highest_levels"], gz, 0.05, True, rs, want_lists=True)
    assert n == int(g["num_rendered"])
    assert np.array_equal(radii.cpu().numpy(), g["radii"])
    assert np.array_equal(pl.cpu().numpy(), g["point_list"])
    assert np.array_equal(rg.cpu().numpy(), g["ranges"])
    assert np.array_equal(color.cpu().numpy(), g["color"])


@pytest.mark.parametrize("level", [0, 1])
def test_mmfr_matches_reference_golden(scene_small, golden_dir, level):
    g = _g(golden_dir, f"mmfr_small_c0_g0_l{level}.npz")
    s, c = scene_small
    sub = {k: (v[:: 1 << level] if isinstance(v, np.ndarray) and v.ndim > 0 and v.shape[0] == s["means3D"].shape[0] else v)
           for k, v in s.items()}
    sc = _cuda({k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in sub.items()})
    import diff_gaussian_rasterization_mmfr_pcheck_obb as m
    rs = _settings(m, c, s["sh_degree"])
    gz = torch.from_numpy(np.asarray(g["gaze"], np.float32)).cuda()
    n, color, radii, pl, rg, item = ops.forward_mmfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                     level, gz, 0.05, True, rs, want_lists=True)
    assert n == int(g["num_rendered"])
    assert np.array_equal(radii.cpu().numpy(), g["radii"])
    assert np.array_equal(pl.cpu().numpy(), g["point_list"])
    assert np.array_equal(rg.cpu().numpy(), g["ranges"])
    assert np.array_equal(color.cpu().numpy(), g["color"])


@pytest.mark.parametrize("gi", [0, 1])
def test_fov_matches_reference_golden(scene_small, golden_dir, gi):
    g = _g(golden_dir, f"fov_small_c0_g{gi}.npz")
    s, c = scene_small
    (n, color, radii, pl, rg, item), _, _ = _run_fov(synth.add_foveation(s), c, g["gaze"])
    assert n == int(g["num_rendered"])
    assert np.array_equal(radii.cpu().numpy(), g["radii"])
    assert np.array_equal(pl.cpu().numpy(), g["point_list"])
    assert np.array_equal(rg.cpu().numpy(), g["ranges"])
    assert np.abs(color.cpu().numpy() - g["color"]).max() <= IMG_TOL
    lvl, mn, gx, gy, bl = ops.fov_tile_tables(item, c["image_width"], c["image_height"])
    # tile tables against the REFERENCE binary's own kernels (decoded by oracle/ref_tile_tables.py when the golden was made)
    for ours, key in ((lvl, "tile_level_ref"), (mn, "tile_min_ref"), (gx, "tile_grad_x_ref"), (gy, "tile_grad_y_ref")):
        assert np.array_equal(ours.cpu().numpy().view(np.int32), g[key].view(np.int32)), key
    assert np.array_equal(bl.cpu().numpy(), g["tile_blend_ref"])


# ---------------------------------------------------------------------------------------------------------------
# 2. CPU oracle on fresh inputs, ragged image sizes, edge cases
# ---------------------------------------------------------------------------------------------------------------
def _small_cam(W, H):
    return synth.look_at_camera(W, H, 70.0, (0.3, 0.2, -3.5))


@pytest.mark.parametrize("W,H,P,seed", [(160, 96, 3000, 5), (250, 130, 4000, 9), (64, 64, 500, 11)])
def test_obb_vs_oracle_ragged_sizes(W, H, P, seed):
    import oracle
    s = synth.make_scene_cube(P, seed)
    c = _small_cam(W, H)
    o = oracle.forward_ps1(s, c, "obb")
    (n, color, radii, item, pl, rg), _, _ = _run_ps1(ops.MODE_OBB, s, c)
    assert n == o["num_rendered"]
    assert np.array_equal(radii.cpu().numpy(), o["radii"])
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
    assert np.array_equal(rg.cpu().numpy().astype(np.uint32), o["ranges"])
    assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL


@pytest.mark.parametrize("gaze", [(0.5, 0.5), (0.1, 0.9), (0.75, 0.25)])
def test_fov_vs_oracle(gaze):
    import oracle
    s = synth.add_foveation(synth.make_scene_cube(5000, 21))
    c = _small_cam(400, 240)
    o = oracle.forward_fov(s, c, gaze)
    (n, color, radii, pl, rg, item), _, _ = _run_fov(s, c, gaze)
    lvl, mn, gx, gy, bl = ops.fov_tile_tables(item, 400, 240)
    # libdevice vs libm acosf/tanf: levels agree to a few ulp; the derived integer decisions must agree
    assert np.abs(mn.cpu().numpy() - o["tile_min"]).max() <= 1e-5
    assert np.array_equal(bl.cpu().numpy(), o["tile_blend"])
    assert n == o["num_rendered"]
    assert np.array_equal(radii.cpu().numpy(), o["radii"])
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
    assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL


@pytest.mark.parametrize("want_lists", [False, True])
def test_uint8_output_is_the_quantised_fp32_image(want_lists):
    """The optional 8-bit image of the foveated blend epilogue (lazy and full-sort paths) is, byte for byte, what the reference's
    scripts would store of the fp32 image: torchvision.utils.save_image = mul(255).add_(0.5).clamp_(0, 255).to(uint8)
    (fov3dgs/render.py:52).  A bright background exercises the upper clamp."""
    import diff_gaussian_rasterization_fov_pcheck_obb as m
    s = synth.add_foveation(synth.make_scene_cube(5000, 23))
    c = _small_cam(400, 240)
    sc = _cuda(s)
    bg = torch.tensor([1.3, 0.4, -0.2], device="cuda")
    rs = _settings(m, c, s["sh_degree"], bg=bg)
    g = torch.tensor([0.4, 0.6], device="cuda")
    args = (sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"], sc["highest_levels"], g,
            0.05, True, rs)
    ref = ops.forward_fov(*args, want_lists=want_lists)
    u8 = ops.forward_fov(*args, want_lists=want_lists, out_uint8=True)
    assert u8[1].dtype == torch.uint8 and u8[0] == ref[0] and torch.equal(u8[2], ref[2])
    expect = ref[1].clone().mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)
    assert torch.equal(u8[1], expect)
    assert int(expect.max()) == 255 and int(expect.min()) == 0
    if not want_lists:
        r = m.GaussianRasterizer(raster_settings=rs)
        r.output_uint8 = True
        color, radii = r(means3D=sc["means3D"], means2D=None, opacities=sc["opacities4"], shs_rest=sc["shs_rest"], scales=sc["scales"],
                         rotations=sc["rotations"], shs_dcs=sc["shs_dcs"], highest_levels=sc["highest_levels"], gazeArray=g,
                         alpha=0.05, blending=True)
        assert torch.equal(color, expect) and torch.equal(radii, ref[2])


@pytest.mark.parametrize("gaze", [(0.5, 0.5), (0.2, 0.8)])
def test_smfr_baseline_vs_oracle_full_and_lazy(gaze):
    """SMFR baseline (naive_pcheck_obb): lists exact, image within tolerance, lazy == full-sort bits;
    its blending tiles follow the shared-alpha rule of naive_pcheck_obb/cuda_rasterizer/forward.cu:383-430."""
    import oracle
    s = synth.add_foveation(synth.make_scene_cube(5000, 23))
    c = _small_cam(400, 240)
    o = oracle.forward_smfr(s, c, gaze)
    sc = _cuda(s)
    import diff_gaussian_rasterization_naive_pcheck_obb as m
    rs = _settings(m, c, s["sh_degree"])
    g = torch.tensor(np.asarray(gaze, np.float32)).cuda()
    n, color, radii, pl, rg, item = ops.forward_smfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                     sc["highest_levels"], g, 0.05, True, rs, want_lists=True)
    assert n == o["num_rendered"]
    assert np.array_equal(radii.cpu().numpy(), o["radii"])
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
    assert np.array_equal(rg.cpu().numpy().astype(np.uint32), o["ranges"])
    assert int(ops.last_stats["num_blend_tiles"]) > 0
    assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL
    r = m.GaussianRasterizer(raster_settings=rs)
    col_l, radii_l = r(means3D=sc["means3D"], means2D=None, opacities=sc["opacity"], shs=sc["shs"], scales=sc["scales"],
                       rotations=sc["rotations"], highest_levels=sc["highest_levels"], gazeArray=g, alpha=0.05, blending=True)
    assert torch.equal(col_l, color) and torch.equal(radii_l, radii)


def test_mmfr_baseline_vs_oracle_and_levels_sum_to_full_image():
    """MMFR baseline (mmfr_pcheck_obb): per level call lists exact, image within tolerance, lazy ==
    full-sort bits; with the SAME model at all four levels the level images add up to the PS=1 image (the blending
    weights of a level pair sum to one, skipped tiles contribute zero)."""
    import oracle
    import diff_gaussian_rasterization_mmfr_pcheck_obb as m
    s = synth.make_scene_cube(5000, 29)
    c = _small_cam(400, 240)
    gaze = (0.35, 0.6)
    sc = _cuda(s)
    rs = _settings(m, c, s["sh_degree"])
    g = torch.tensor(np.asarray(gaze, np.float32)).cuda()
    total = torch.zeros((3, 240, 400), device="cuda")
    r = m.GaussianRasterizer(raster_settings=rs)
    for level in range(4):
        o = oracle.forward_mmfr(s, c, level, gaze)
        n, color, radii, pl, rg, item = ops.forward_mmfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                         level, g, 0.05, True, rs, want_lists=True)
        assert n == o["num_rendered"], level
        assert np.array_equal(radii.cpu().numpy(), o["radii"])
        assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
        assert np.array_equal(rg.cpu().numpy().astype(np.uint32), o["ranges"])
        assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL
        col_l, radii_l = r(means3D=sc["means3D"], means2D=None, opacities=sc["opacity"], shs=sc["shs"], scales=sc["scales"],
                           rotations=sc["rotations"], cur_level=level, gazeArray=g, alpha=0.05, blending=True)
        assert torch.equal(col_l, color) and torch.equal(radii_l, radii)
        total += color
    (n1, full, _, _, _, _), _, _ = _run_ps1(ops.MODE_OBB, s, c)
    assert float((total - full).abs().max()) <= 1e-5


def test_render_from_ply_and_camera_json(tmp_path):
    """north_star: "identical point_cloud.ply + camera JSON inputs" — the whole input path (fovgs.io: PLY files of the four
    levels, compose rule, cameras.json) feeding the foveated rasterizer, checked against the oracle on the same files."""
    import json
    import oracle
    from fovgs import io
    rng = np.random.default_rng(8)
    s = synth.make_scene_cube(4000, 8)
    P = s["means3D"].shape[0]
    raw0 = {"xyz": s["means3D"], "features_dc": s["shs"][:, :1], "features_rest": s["shs"][:, 1:],
            "opacity": np.log(s["opacity"] / (1 - s["opacity"])).astype(np.float32), "scaling": np.log(s["scales"]).astype(np.float32),
            "rotation": s["rotations"], "sh_degree": 3}
    paths = [str(tmp_path / "l0" / "point_cloud.ply")]
    io.write_ply(paths[0], raw0)
    keep = np.arange(P)
    for i in range(1, 4):
        keep = np.sort(rng.choice(keep, size=len(keep) // 2, replace=False))
        raw = {k: (v[keep] if isinstance(v, np.ndarray) else v) for k, v in raw0.items()}
        raw["features_dc"] = raw["features_dc"] + rng.normal(0, 0.2, raw["features_dc"].shape).astype(np.float32)
        raw["indexes"] = keep.astype(np.int32).reshape(-1, 1)
        paths.append(str(tmp_path / f"l{i}" / "point_cloud.ply"))
        io.write_ply(paths[-1], raw, with_index=True)
    c0 = _small_cam(320, 208)
    wv = c0["viewmatrix"].astype(np.float64)
    entry = io.camera_to_json_entry(0, wv[:3, :3], wv[3, :3], c0["FoVx"], c0["FoVy"], 320, 208, "view0")
    cj = str(tmp_path / "cameras.json")
    json.dump([entry], open(cj, "w"))
    scene = io.compose_levels([io.raw_model_from_ply(p) for p in paths])
    cam = io.cameras_from_json(cj)[0]
    assert scene["highest_levels"].max() == 3 and cam["image_width"] == 320
    o = oracle.forward_fov(scene, cam, (0.4, 0.5))
    (n, color, radii, pl, rg, item), _, _ = _run_fov(scene, cam, (0.4, 0.5))
    assert n == o["num_rendered"] and np.array_equal(radii.cpu().numpy(), o["radii"])
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
    assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL


def test_large_image_uses_global_level_table():
    """More tiles than the shared level-code table of k_pre holds (LC_MAX = 16384): the per-candidate level test falls back
    to the global float table.  2608 x 2000 -> 163 x 125 = 20375 tiles; foveated, SMFR and MMFR against the oracle."""
    import oracle
    W, H = 2608, 2000
    s = synth.add_foveation(synth.make_scene_cube(3000, 51))
    s["scales"] = (s["scales"] * 2.0).astype(np.float32)
    c = synth.look_at_camera(W, H, 65.0, (0.2, 0.1, -3.2))
    gaze = (0.45, 0.55)
    o = oracle.forward_fov(s, c, gaze, list_cap=1 << 22)
    (n, color, radii, pl, rg, item), sc, rs = _run_fov(s, c, gaze)
    assert n == o["num_rendered"] and np.array_equal(radii.cpu().numpy(), o["radii"])
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
    assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL
    g = torch.tensor(np.asarray(gaze, np.float32)).cuda()
    o2 = oracle.forward_smfr(s, c, gaze)
    n2, col2, rad2 = ops.forward_smfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"], sc["highest_levels"],
                                      g, 0.05, True, rs)
    assert n2 == o2["num_rendered"] and np.abs(col2.cpu().numpy() - o2["color"]).max() <= IMG_TOL
    o3 = oracle.forward_mmfr(s, c, 1, gaze)
    n3, col3, rad3 = ops.forward_mmfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"], 1, g, 0.05, True, rs)
    assert n3 == o3["num_rendered"] and np.array_equal(rad3.cpu().numpy(), o3["radii"])
    assert np.abs(col3.cpu().numpy() - o3["color"]).max() <= IMG_TOL


def test_conservative_cull_is_exact_on_adversarial_inputs():
    """k_pre phase 0 drops Gaussians whose tile rectangle must be empty.  Stress it where a loose bound would show: huge and
    strongly anisotropic splats just outside the screen, unnormalised quaternions, Gaussians hugging the near plane, a scaled
    (non-rigid) view matrix, scale_modifier != 1 — lists and radii must still equal the oracle's exactly."""
    import oracle
    import diff_gaussian_rasterization_pcheck_obb as m
    rng = np.random.default_rng(77)
    P = 6000
    s = synth.make_scene_cube(P, 77)
    s["means3D"] = (rng.uniform(-1, 1, (P, 3)) * np.array([6.0, 6.0, 1.5])).astype(np.float32)     # most are off-screen
    s["scales"] = np.exp(rng.normal(-1.5, 1.3, (P, 3))).astype(np.float32)                          # some cover the whole image
    q = rng.normal(size=(P, 4)) * rng.uniform(0.5, 1.6, (P, 1))                                      # |q| != 1
    s["rotations"] = q.astype(np.float32)
    c = synth.look_at_camera(304, 176, 55.0, (0.0, 0.0, -2.2))
    near = rng.choice(P, 300, replace=False)                                                         # hug the near plane z = 0.2
    s["means3D"][near, 2] = (-2.2 + 0.2 + rng.normal(0, 0.01, 300)).astype(np.float32)
    o = oracle.forward_ps1(s, c, "obb")
    (n, color, radii, item, pl, rg), _, _ = _run_ps1(ops.MODE_OBB, s, c)
    assert n == o["num_rendered"] and np.array_equal(radii.cpu().numpy(), o["radii"])
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])
    assert (o["radii"] > 0).sum() > 200 and (o["radii"] == 0).sum() > 2000


@pytest.mark.parametrize("kind,P", [("uniform", 20000), ("clustered", 30000), ("tiny", 3), ("duplicates", 5000)])
def test_simple_knn_distCUDA2_matches_exact_knn(kind, P):
    """simple_knn._C.distCUDA2: mean squared distance to the 3 nearest neighbours, against an exact
    k-d tree on the CPU (scipy), on uniform, strongly clustered (COLMAP-like) and degenerate inputs."""
    from scipy.spatial import cKDTree
    from simple_knn._C import distCUDA2
    rng = np.random.default_rng(len(kind) + P)
    if kind == "uniform":
        pts = rng.uniform(-5, 5, (P, 3))
    elif kind == "clustered":
        centers = rng.normal(0, 10, (40, 3))
        pts = centers[rng.integers(0, 40, P)] + rng.normal(0, 0.05, (P, 3)) * rng.uniform(0.1, 3.0, (P, 1))
        pts[:50] = rng.uniform(-200, 200, (50, 3))            # far outliers stretch the grid
    elif kind == "duplicates":
        pts = np.repeat(rng.uniform(-1, 1, (P // 5, 3)), 5, axis=0)
    else:
        pts = rng.uniform(-1, 1, (P, 3))
    pts = pts.astype(np.float32)
    out = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    if P < 4:
        assert out.shape == (P,) and (out > 1e37).all()               # fewer than 3 neighbours: FLT_MAX-based, like the reference
        return
    d, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=4)
    ref = (d[:, 1:] ** 2).mean(axis=1)
    assert np.allclose(out, ref, rtol=2e-5, atol=1e-12)


def test_sum_backward_vs_oracle():
    import oracle
    s = synth.make_scene_cube(3000, 33)
    c = _small_cam(176, 112)
    o = oracle.forward_ps1(s, c, "sum")
    (n, color, radii, item, gcount, contrib, pl, rg), sc, rs = _run_ps1(ops.MODE_SUM, s, c)
    assert np.array_equal(gcount.cpu().numpy(), o["gaussians_count"])
    grad_out = np.random.default_rng(1).standard_normal((3, 112, 176)).astype(np.float32)
    go = oracle.backward_ps1(s, c, o, grad_out)
    g = ops.backward_ps1(item, sc["means3D"], radii, sc["scales"], sc["rotations"], None, sc["shs"], None, rs,
                         torch.from_numpy(grad_out).cuda())
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for nm, t in zip(names, g):
        a, b = t.cpu().numpy().ravel(), go[nm].ravel()
        assert np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30) <= 1e-4, nm


@pytest.mark.parametrize("variant", ["sum", "max", "lwmc"])
def test_training_family_statistics_vs_oracle_full_and_lazy(variant):
    """SUM / MAX / LWMC: per-Gaussian statistics against the CPU oracle, through both the full-sort
    kernel (per-hit atomics, like the reference) and the lazy kernel (warp/tile-level reduction).  Counts are integers
    -> exact; MAX's contribution is an exact maximum -> equal up to libm-vs-libdevice expf (1e-5 relative); SUM/LWMC are
    fp32 sums in arbitrary order -> 1e-4 relative, and a pixel's arg-max may flip on an expf ulp -> a handful of
    Gaussians may trade one pixel's loss, the total is conserved."""
    import oracle
    s = synth.make_scene_cube(6000, 41)
    W, H = 208, 144
    c = _small_cam(W, H)
    mode = {"sum": ops.MODE_SUM, "max": ops.MODE_MAX, "lwmc": ops.MODE_LWMC}[variant]
    lm = np.random.default_rng(5).random((H, W)).astype(np.float32) if variant == "lwmc" else None
    o = oracle.forward_ps1(s, c, variant, loss_map=lm)
    sc = _cuda(s)
    import diff_gaussian_rasterization_pcheck_obb_sum as m
    rs = _settings(m, c, s["sh_degree"])
    lmt = None if lm is None else torch.from_numpy(lm).cuda()
    outs = {}
    for lazy in (False, True):
        r = ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs,
                            want_lists=not lazy, loss_map=lmt)
        n, color, radii, item, gcount, contrib = r[:6]
        outs[lazy] = (color, gcount, contrib, item, radii)
        assert n == o["num_rendered"]
        assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL
        assert np.array_equal(gcount.cpu().numpy(), o["gaussians_count"]), f"{variant} lazy={lazy}"
        a, b = contrib.cpu().numpy().astype(np.float64), o["contributions"].astype(np.float64)
        if variant == "max":
            assert np.abs(a - b).max() <= 1e-5
        elif variant == "sum":
            assert (np.abs(a - b) / (np.abs(b) + 1e-3)).max() <= 1e-4
        else:
            assert abs(a.sum() - float(lm.sum())) <= 1e-3 * lm.sum()      # every inside pixel gave its loss to someone
            assert abs(a[0] - b[0]) <= 1e-3 * max(1.0, b[0])              # Gaussian 0 collects the empty pixels
            assert (np.abs(a - b) > 1e-3 * (1.0 + np.abs(b))).mean() <= 1e-3
    # lazy == full: same image bits, same counts, same saved state for backward
    assert torch.equal(outs[False][0], outs[True][0]) and torch.equal(outs[False][1], outs[True][1])
    if variant == "max":
        assert torch.equal(outs[False][2], outs[True][2])
    else:
        assert ((outs[False][2] - outs[True][2]).abs() / (outs[False][2].abs() + 1e-3)).max() <= 1e-4
    grad_out = torch.from_numpy(np.random.default_rng(2).standard_normal((3, H, W)).astype(np.float32)).cuda()
    g = [ops.backward_ps1(outs[z][3], sc["means3D"], outs[z][4], sc["scales"], sc["rotations"], None, sc["shs"], None, rs, grad_out)
         for z in (False, True)]
    for a, b in zip(*g):
        assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12


def test_loss_weighted_drop_in_package_takes_loss_map(scene_small):
    import diff_gaussian_rasterization_pcheck_obb_loss_weighted_max_count as m
    import diff_gaussian_rasterization_pcheck_obb_max as mm
    s, c = scene_small
    sc = _cuda(s)
    H, W = c["image_height"], c["image_width"]
    lm = torch.rand((H, W), device="cuda")
    r = m.GaussianRasterizer(raster_settings=_settings(m, c, 3))
    color, radii, cnt, contrib = r(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacity"], shs=sc["shs"],
                                   scales=sc["scales"], rotations=sc["rotations"], loss_map=lm)
    assert abs(float(contrib.sum()) - float(lm.sum())) <= 1e-3 * float(lm.sum())
    with pytest.raises(RuntimeError, match="loss_map"):
        r(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacity"], shs=sc["shs"], scales=sc["scales"],
          rotations=sc["rotations"])
    r2 = mm.GaussianRasterizer(raster_settings=_settings(mm, c, 3))
    color2, radii2, cnt2, contrib2 = r2(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacity"], shs=sc["shs"],
                                        scales=sc["scales"], rotations=sc["rotations"])
    assert torch.equal(color, color2) and float(contrib2.max()) <= 0.99 and int(cnt2.sum()) > int(cnt.sum())


def test_edge_cases_empty_culled_single_and_huge():
    import diff_gaussian_rasterization_pcheck_obb as m
    c = synth.config1_camera()
    s = synth.make_scene_cube(8, 0)
    rs = _settings(m, c, 3)
    r = m.GaussianRasterizer(raster_settings=rs)
    # empty scene: image stays zero, like the reference (rasterize_points.cu: P == 0 skips the kernel)
    e = torch.zeros((0, 3), device="cuda")
    color, radii = r(means3D=e, means2D=e, opacities=torch.zeros((0, 1), device="cuda"), shs=torch.zeros((0, 16, 3), device="cuda"),
                     scales=e, rotations=torch.zeros((0, 4), device="cuda"))
    assert color.shape == (3, 256, 256) and float(color.abs().max()) == 0.0 and radii.numel() == 0
    # everything behind the camera
    sc = _cuda(s)
    behind = sc["means3D"].clone()
    behind[:, 2] = -100.0
    color, radii = r(means3D=behind, means2D=behind, opacities=sc["opacity"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    assert int(radii.abs().sum()) == 0 and float(color.abs().max()) == 0.0
    # one splat that covers every tile (exercises the multi-round candidate expansion)
    import oracle
    big = {k: (v[:1].copy() if isinstance(v, np.ndarray) else v) for k, v in s.items()}
    big["means3D"][:] = 0.0
    big["scales"][:] = 3.0
    big["opacity"][:] = 0.9
    o = oracle.forward_ps1(big, c, "obb")
    (n, col, rad, item, pl, rg), _, _ = _run_ps1(ops.MODE_OBB, big, c)
    assert n == o["num_rendered"] and n >= 1
    assert np.abs(col.cpu().numpy() - o["color"]).max() <= IMG_TOL


def test_prefiltered_flag_reports_near_plane_violations(scene_small):
    """`prefiltered=True` promises that no Gaussian is behind the near plane; the reference prints and __trap()s when the
    promise is broken (auxiliary.h:286-293).  Here the call raises RuntimeError with the reference's message instead of
    killing the CUDA context; with the promise kept the flag changes nothing."""
    import diff_gaussian_rasterization_pcheck_obb as m
    s, c = scene_small
    sc = _cuda(s)
    cc = _cuda(c)

    def rs(pref):
        return m.GaussianRasterizationSettings(c["image_height"], c["image_width"], c["tanfovx"], c["tanfovy"], torch.zeros(3, device="cuda"),
                                               1.0, cc["viewmatrix"], cc["projmatrix"], 3, cc["campos"], pref, False)

    kw = dict(means2D=None, opacities=sc["opacity"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    a = m.GaussianRasterizer(raster_settings=rs(False))(means3D=sc["means3D"], **kw)[0]
    b = m.GaussianRasterizer(raster_settings=rs(True))(means3D=sc["means3D"], **kw)[0]      # scene is entirely in front
    assert torch.equal(a, b)
    behind = sc["means3D"].clone()
    behind[:7, 2] = -50.0
    with pytest.raises(RuntimeError, match="filtered although prefiltered is set"):
        m.GaussianRasterizer(raster_settings=rs(True))(means3D=behind, **kw)
    m.GaussianRasterizer(raster_settings=rs(False))(means3D=behind, **kw)                    # without the promise: just culled


def test_equal_depth_ties_are_ordered_by_id():
    """Many Gaussians on one fronto-parallel plane: equal depth bits -> the stable order is by Gaussian id."""
    import oracle
    s = synth.make_scene_cube(2000, 3)
    s["means3D"][:, 2] = 0.25   # same depth for all
    c = synth.config1_camera()
    o = oracle.forward_ps1(s, c, "obb")
    (n, color, radii, item, pl, rg), _, _ = _run_ps1(ops.MODE_OBB, s, c)
    assert np.array_equal(pl.cpu().numpy().astype(np.uint32), o["point_list"])


def test_lazy_sort_blend_equals_full_sort_blend():
    """The consumption-driven (lazy) path and the full-sort path must produce bit-identical images, including the
    oversize-bucket fallback (thousands of instances per tile with identical depth bits) and multi-bucket tiles."""
    import oracle
    # (a) dense fronto-parallel slab: every tile holds > 2048 instances with equal depth
    s = synth.make_scene_cube(30000, 4)
    s["means3D"][:, 2] = 0.5
    s["scales"][:] *= 3.0
    c = synth.look_at_camera(64, 64, 50.0, (0.0, 0.0, -3.0))
    (n, col_full, radii, item, pl, rg), _, _ = _run_ps1(ops.MODE_OBB, s, c, want_lists=True)
    assert int(ops.last_stats["max_tile_instances"]) > 2048
    (n2, col_lazy, radii2, item2), _, _ = _run_ps1(ops.MODE_OBB, s, c, want_lists=False)
    assert n2 == n and torch.equal(col_full, col_lazy)
    o = oracle.forward_ps1(s, c, "obb")
    assert np.abs(col_lazy.cpu().numpy() - o["color"]).max() <= IMG_TOL
    # (b) deep scene, many depth buckets per tile, foveated + blending tiles
    f = synth.add_foveation(synth.make_scene_bicycle(400000, 2, log_scale_mu=-3.2))
    cam = synth.ring_cameras(30, 640, 360)[5]
    (nf, col_f, _, _, _, _), _, _ = _run_fov(f, cam, (0.5, 0.5), want_lists=True)
    assert int(ops.last_stats["max_tile_instances"]) > 2048
    (nl, col_l, _), _, _ = _run_fov(f, cam, (0.5, 0.5), want_lists=False)
    assert nl == nf and torch.equal(col_f, col_l)
    ops.set_full_sort(True)
    try:
        (nl2, col_l2, _), _, _ = _run_fov(f, cam, (0.5, 0.5), want_lists=False)
    finally:
        ops.set_full_sort(False)
    assert torch.equal(col_f, col_l2)


def test_tma_colour_stage_equals_register_staged_stage(scene_small):
    """k_color_tma (cp.async.bulk gathers) and k_color (register-staged loads) must give identical colours, including
    Gaussian 0 / P-1 (window clipping at the tensor ends) and odd ids (4-byte aligned SH blocks)."""
    s, c = scene_small
    f = synth.add_foveation(s)
    outs = []
    for no_tma in (False, True):
        ops.set_no_tma(no_tma)
        try:
            (n, col, radii, pl, rg, item), _, _ = _run_fov(f, c, (0.5, 0.5))
            geo = ops.geometry(item, ops.MODE_FOV, radii.numel(), c["image_width"], c["image_height"])
            (n2, col2, radii2, item2, pl2, rg2), _, _ = _run_ps1(ops.MODE_OBB, s, c)
            geo2 = ops.geometry(item2, ops.MODE_OBB, radii2.numel(), c["image_width"], c["image_height"])
            vis = (radii > 0).unsqueeze(-1).unsqueeze(-1)
            outs.append((col.clone(), col2.clone(), (geo["level_colors"] * 1.0).clone(), geo2["rgb"].clone(), radii.clone(), radii2.clone()))
        finally:
            ops.set_no_tma(False)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    v2 = outs[0][5] > 0
    assert torch.equal(outs[0][3][v2], outs[1][3][v2])


def test_packed_model_cache_is_bit_identical_and_tracks_updates(scene_small):
    """The foveated colour stage reads the static model tensors through a packed copy (one aligned 256-byte row per
    Gaussian) cached on the caller's tensor objects.  Same bits with and without it; an in-place update (version counter)
    re-packs; `.data` writes are caught in debug mode; dead tensors drop their entry."""
    import gc
    import diff_gaussian_rasterization_fov_pcheck_obb as m
    s, c = scene_small
    sc = _cuda(synth.add_foveation(s))
    rs = _settings(m, c, 3)
    g = torch.tensor([0.4, 0.6], device="cuda")

    def run(settings=rs):
        return ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"],
                               sc["highest_levels"], g, 0.05, True, settings)[1].clone()

    ops.invalidate_model_cache()
    ops.set_model_cache(False)
    ref = run()
    ops.set_model_cache(True)
    a0 = run()                                    # first sighting of these tensor objects: un-packed gather, nothing cached
    assert len(ops._packed_cache) == 0
    a = run()                                     # second consecutive call with the same objects: packed
    assert len(ops._packed_cache) == 1
    b = run()
    assert len(ops._packed_cache) == 1 and torch.equal(ref, a0) and torch.equal(ref, a) and torch.equal(ref, b)
    sc["shs_dcs"].mul_(0.5)                       # in-place: version bump -> new key -> re-pack (at the second call)
    ops.set_model_cache(False); ref2 = run(); ops.set_model_cache(True)
    run()
    c2 = run()
    assert torch.equal(ref2, c2) and not torch.equal(ref, c2) and len(ops._packed_cache) == 1
    sc["opacities4"].data.mul_(0.5)               # bypasses the version counter: stale rows, caught by debug mode
    with pytest.raises(RuntimeError, match="stale"):
        run(_settings(m, c, 3, debug=True))
    ops.invalidate_model_cache()
    run()
    assert torch.equal(run(), run(_settings(m, c, 3, debug=True)))
    n_before = len(ops._packed_cache)
    sc["shs_dcs"] = sc["shs_dcs"].clone()         # the old tensor object dies -> its entry goes with it
    gc.collect()
    assert len(ops._packed_cache) == n_before - 1


@pytest.mark.parametrize("deg", [0, 1, 2])
def test_lower_sh_degrees_through_packed_and_unpacked_colour_paths(deg):
    """Models with fewer SH coefficients (M_rest = 0, 3, 8): the packed colour rows are zero padded, the TMA windows shrink."""
    import oracle
    s = synth.make_scene_cube(3000, 61)
    s["shs"] = np.ascontiguousarray(s["shs"][:, : (deg + 1) ** 2])
    s["sh_degree"] = deg
    f = synth.add_foveation(s)
    c = _small_cam(256, 160)
    o = oracle.forward_fov(f, c, (0.5, 0.4))
    for cache in (True, False):
        ops.set_model_cache(cache)
        try:
            (n, color, radii, pl, rg, item), _, _ = _run_fov(f, c, (0.5, 0.4))
        finally:
            ops.set_model_cache(True)
        assert n == o["num_rendered"] and np.array_equal(radii.cpu().numpy(), o["radii"])
        assert np.abs(color.cpu().numpy() - o["color"]).max() <= IMG_TOL
    o1 = oracle.forward_ps1(s, c, "obb")
    (n1, col1, rad1, item1, pl1, rg1), _, _ = _run_ps1(ops.MODE_OBB, s, c)
    assert n1 == o1["num_rendered"] and np.abs(col1.cpu().numpy() - o1["color"]).max() <= IMG_TOL


def test_capacity_overflow_regrows(monkeypatch):
    monkeypatch.setenv("FOVGS_INSTANCE_CAPACITY", "1000")
    ops._pool.clear()
    s = synth.make_scene_cube(10000, 0)
    (n, color, radii, item, pl, rg), _, _ = _run_ps1(ops.MODE_OBB, s, synth.config1_camera())
    assert n == 57970 and item["cap"] >= n
    ops._pool.clear()


def test_drop_in_classes_and_mark_visible(scene_small):
    import diff_gaussian_rasterization_pcheck_obb_sum as msum
    s, c = scene_small
    sc = _cuda(s)
    rs = _settings(msum, c, 3)
    r = msum.GaussianRasterizer(raster_settings=rs)
    means = sc["means3D"].clone().requires_grad_(True)
    color, radii, cnt, contrib = r(means3D=means, means2D=torch.zeros_like(means), opacities=sc["opacity"], shs=sc["shs"],
                                   scales=sc["scales"], rotations=sc["rotations"])
    color.sum().backward()
    assert means.grad is not None and torch.isfinite(means.grad).all() and float(means.grad.abs().sum()) > 0
    vis = r.markVisible(sc["means3D"])
    z = (torch.cat([sc["means3D"], torch.ones_like(sc["means3D"][:, :1])], 1) @ _cuda(c)["viewmatrix"])[:, 2]
    assert vis.dtype == torch.bool and bool((vis == (z > 0.2)).all())


def test_debug_mode_matches_async_mode(scene_small):
    import diff_gaussian_rasterization_pcheck_obb as m
    s, c = scene_small
    sc = _cuda(s)
    outs = []
    for dbg in (False, True):
        r = m.GaussianRasterizer(raster_settings=_settings(m, c, 3, debug=dbg))
        outs.append(r(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacity"], shs=sc["shs"],
                      scales=sc["scales"], rotations=sc["rotations"])[0])
    assert torch.equal(outs[0], outs[1])


# ---------------------------------------------------------------------------------------------------------------
# 3. full-size properties (BASELINE.json configs 2/3: 6 M Gaussians, 1920x1080)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big_scene():
    return synth.add_foveation(synth.make_scene_bicycle(6_000_000, 1)), synth.ring_cameras(30)[0]


def test_fullsize_foveated_properties(big_scene):
    s, c = big_scene
    (n, color, radii, pl, rg, item), sc, _ = _run_fov(s, c, (0.5, 0.5))
    T = 120 * 68
    rg = rg.cpu().numpy().astype(np.int64)
    pl_t = pl.long()
    assert n == pl.numel() == int(rg[:, 1].max())
    nonempty = rg[:, 1] > rg[:, 0]
    # ranges tile the list exactly once, in tile order
    starts, ends = rg[nonempty, 0], rg[nonempty, 1]
    assert starts[0] == 0 and np.array_equal(starts[1:], ends[:-1]) and ends[-1] == n
    # every listed Gaussian is visible, every visible Gaussian is listed
    listed = torch.zeros(radii.numel(), dtype=torch.bool, device="cuda")
    listed[pl_t] = True
    assert bool((listed == (radii > 0)).all())
    # per-tile order: depth non-decreasing, ties by id
    geo = ops.geometry(item, ops.MODE_FOV, radii.numel(), 1920, 1080)
    d = geo["depths"][pl_t]
    tile_of = torch.repeat_interleave(torch.arange(T, device="cuda"), torch.from_numpy(rg[:, 1] - rg[:, 0]).cuda())
    same_tile = tile_of[1:] == tile_of[:-1]
    assert bool(((d[1:] >= d[:-1]) | ~same_tile).all())
    ties = same_tile & (d[1:] == d[:-1])
    assert bool(((pl_t[1:] > pl_t[:-1]) | ~ties).all())
    assert bool(torch.isfinite(color).all()) and float(color.min()) >= 0.0
    # idempotence / determinism: the second frame is bit-identical (no atomics on the value path)
    (n2, color2, radii2, pl2, rg2, _), _, _ = _run_fov(s, c, (0.5, 0.5))
    assert n2 == n and torch.equal(color, color2) and torch.equal(pl, pl2) and torch.equal(radii, radii2)


def test_fullsize_gaze_changes_only_the_foveated_work(big_scene):
    """Moving the gaze changes binning (levels) but never the projection: radius of a kept Gaussian is gaze-free."""
    s, c = big_scene
    (n1, _, radii1, _, _, _), _, _ = _run_fov(s, c, (0.25, 0.25), want_lists=True)
    (n2, _, radii2, _, _, _), _, _ = _run_fov(s, c, (0.75, 0.75), want_lists=True)
    both = (radii1 > 0) & (radii2 > 0)
    assert n1 != n2 and bool((radii1[both] == radii2[both]).all())


def test_integration_binding_stub_renders_like_the_package(scene_small):
    """The ctypes stub INTEGRATION.md quotes (examples/binding_stub.py) is executed as written: same image, radii and instance
    count as the drop-in package on the same inputs."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("binding_stub", os.path.join(root, "examples", "binding_stub.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    s, c = scene_small
    f = synth.add_foveation(s)
    (n, color, radii, pl, rg, item), sc, rs = _run_fov(f, c, (0.4, 0.6))
    cd = _cuda(c)
    g = torch.tensor([0.4, 0.6], dtype=torch.float32, device="cuda")
    e = torch.empty(0, device="cuda")
    n2, color2, radii2 = stub.rasterize_gaussians_fov(sc["shs_dcs"], sc["highest_levels"], g, 0.05, True, torch.zeros(3, device="cuda"),
                                                      sc["means3D"], e, sc["opacities4"], sc["scales"], sc["rotations"], 1.0, e,
                                                      cd["viewmatrix"], cd["projmatrix"], c["tanfovx"], c["tanfovy"], c["image_height"],
                                                      c["image_width"], sc["shs_rest"], 3, cd["campos"], False, False)
    assert n2 == n and torch.equal(radii2, radii) and torch.equal(color2, color)


def test_blend_exp_is_bit_identical_to_expf():
    """The blend kernels' exp (libdevice expf's sequence with its constants hoisted into registers) against expf itself: every
    float of the blend's domain [-4.5, -0.0] (1.08e9 bit patterns), the positive range up to 4.5, and the whole finite negative /
    positive ranges in 2^24-pattern windows."""
    from fovgs._lib import lib
    L = lib()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ranges = [(0x80000000, 0xC0900000), (0x00000000, 0x40900000)]
    ranges += [(b, b + (1 << 24) - 1) for b in range(0xC0900000, 0xFF800000 - (1 << 24), 1 << 27)]
    ranges += [(b, b + (1 << 24) - 1) for b in range(0x40900000, 0x7F800000 - (1 << 24), 1 << 27)]
    for lo, hi in ranges:
        assert L.fovgs_debug_expf_mismatches(lo, hi, bad.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
