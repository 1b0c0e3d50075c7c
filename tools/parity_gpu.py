#!/usr/bin/env python
"""GPU-box parity + timing report: our sm_100a library vs the UNMODIFIED reference CUDA (oracle/_ref/*.so), on the
same seeded inputs.  Writes gpurun_out/parity_report.json (+ golden fixtures under gpurun_out/golden/ with --golden).

  python tools/parity_gpu.py --sizes small,mid --variants fov,obb,sum [--golden] [--time]
"""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from fovgs import ops, synth  # noqa: E402
ops.set_full_stats(True)   # these tools print the blend stage's counters: wait for the end of each frame
import ref_api  # noqa: E402


def to_cuda(d):
    out = {}
    for k, v in d.items():
        out[k] = torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v
    return out


class RS:  # GaussianRasterizationSettings stand-in
    pass


def settings(cam, sh_degree, bg, debug=False):
    rs = RS()
    rs.image_height = cam["image_height"]; rs.image_width = cam["image_width"]
    rs.tanfovx = cam["tanfovx"]; rs.tanfovy = cam["tanfovy"]
    rs.bg = bg; rs.scale_modifier = 1.0
    rs.viewmatrix = cam["viewmatrix"]; rs.projmatrix = cam["projmatrix"]
    rs.sh_degree = sh_degree; rs.campos = cam["campos"]; rs.prefiltered = False; rs.debug = debug
    return rs


def bits_equal(a, b):
    return bool(torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)))


def mism(a, b, mask=None):
    ai = a.contiguous().view(torch.int32)
    bi = b.contiguous().view(torch.int32)
    ne = ai != bi
    if mask is not None:
        m = mask
        while m.dim() < ne.dim():
            m = m.unsqueeze(-1)
        ne = ne & m
    return int(ne.sum().item())


def compare_common(rep, ref, ours, geo_ref, geo_ours, W, H):
    radii_r, radii_o = ref["radii"], ours["radii"]
    rep["num_rendered_ref"] = int(ref["n"]); rep["num_rendered_ours"] = int(ours["n"])
    rep["radii_mismatch"] = int((radii_r != radii_o).sum().item())
    vis = (radii_r > 0) & (radii_o > 0)
    rep["visible_ref"] = int((radii_r > 0).sum().item())
    for k in ("means2D", "depths", "conic"):
        rep[k + "_bit_mismatch"] = mism(geo_ref[k], geo_ours[k], vis)
    N = int(ref["n"])
    if int(ours["n"]) == N:
        rep["point_list_mismatch"] = int((ref["point_list"][:N] != ours["point_list"][:N]).sum().item())
    else:
        rep["point_list_mismatch"] = -1
    T = ((W + 15) // 16) * ((H + 15) // 16)
    rep["ranges_mismatch"] = int((ref["ranges"][:T] != ours["ranges"][:T]).sum().item())
    d = (ref["color"] - ours["color"]).abs()
    rep["img_max_abs"] = float(d.max().item())
    rep["img_mean_abs"] = float(d.mean().item())
    rep["img_n_gt_1e-4"] = int((d > 1e-4).sum().item())
    rep["img_n_gt_1e-6"] = int((d > 1e-6).sum().item())


def time_fn(fn, warm=5, iters=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return {"ms_median": ts[len(ts) // 2], "ms_min": ts[0], "ms_max": ts[-1]}


def run_fov(scn, cam, gazes, rep_list, do_time, golden_dir, tag):
    mod = ref_api.ref_module("ref_fov_C")
    sc = to_cuda(synth.add_foveation(scn))
    c = to_cuda(cam)
    W, H = cam["image_width"], cam["image_height"]
    P = sc["means3D"].shape[0]
    bg = torch.zeros(3, device="cuda")
    rs = settings(c, sc["sh_degree"], bg)
    for gi, g in enumerate(gazes):
        rep = {"variant": "fov", "tag": tag, "P": P, "W": W, "H": H, "gaze": list(g)}
        try:
            gaze = torch.tensor(g, dtype=torch.float32, device="cuda")
            n_r, col_r, rad_r, geom, binning, img = ref_api.fov_forward(mod, sc, c, gaze)
            torch.cuda.synchronize()
            gr = ref_api.decode_geom(geom, P, "fov")
            br = ref_api.decode_binning(binning, n_r)
            ir = ref_api.decode_img(img, W, H)
            ref = {"n": n_r, "color": col_r, "radii": rad_r, "point_list": br["point_list"], "ranges": ir["ranges"]}
            n_o, col_o, rad_o, pl_o, rg_o, item = ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"],
                                                                  sc["shs_rest"], sc["shs_dcs"], sc["highest_levels"], gaze, 0.05,
                                                                  True, rs, want_lists=True)
            torch.cuda.synchronize()
            go = ops.geometry(item, ops.MODE_FOV, P, W, H)
            ours = {"n": n_o, "color": col_o, "radii": rad_o, "point_list": pl_o, "ranges": rg_o}
            compare_common(rep, ref, ours, gr, go, W, H)
            rep["stats"] = dict(ops.last_stats)
            _, col_l, rad_l = ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"],
                                              sc["highest_levels"], gaze, 0.05, True, rs)
            rep["lazy_img_max_abs"] = float((col_l - col_r).abs().max().item()); rep["lazy_stats"] = dict(ops.last_stats)
            if golden_dir and gi < 2:
                # per-tile tables from the REFERENCE binary's own kernels (oracle/ref_tile_tables.py), not from ours
                import ref_tile_tables
                tt = ref_tile_tables.reference_tile_tables(W, H, g, 0.05)
                np.savez_compressed(os.path.join(golden_dir, f"fov_{tag}_g{gi}.npz"), gaze=np.array(g, np.float32),
                                    color=col_r.cpu().numpy(), radii=rad_r.cpu().numpy(), num_rendered=np.int64(n_r),
                                    point_list=br["point_list"].cpu().numpy(), ranges=ir["ranges"][: ((W + 15) // 16) * ((H + 15) // 16)].cpu().numpy(),
                                    tile_level_ref=tt["tile_level"], tile_min_ref=tt["tile_min"], tile_grad_x_ref=tt["grad_x"],
                                    tile_grad_y_ref=tt["grad_y"], tile_blend_ref=tt["blending"].astype(np.uint8))
            if do_time and gi == 0:
                rep["time_ref"] = time_fn(lambda: ref_api.fov_forward(mod, sc, c, gaze))
                rep["time_ours"] = time_fn(lambda: ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"],
                                                                   sc["shs_rest"], sc["shs_dcs"], sc["highest_levels"], gaze, 0.05, True, rs))
        except Exception as ex:  # keep going: one report per case
            rep["error"] = repr(ex); rep["trace"] = traceback.format_exc()[-1500:]
        rep_list.append(rep)
        print(json.dumps({k: v for k, v in rep.items() if k != "trace"}), flush=True)


def run_smfr(scn, cam, gazes, rep_list, do_time, golden_dir, tag):
    """SMFR baseline (naive_pcheck_obb) vs ref_naive_C."""
    mod = ref_api.ref_module("ref_naive_C")
    sc = to_cuda(synth.add_foveation(scn))
    c = to_cuda(cam)
    W, H = cam["image_width"], cam["image_height"]
    P = sc["means3D"].shape[0]
    bg = torch.zeros(3, device="cuda")
    rs = settings(c, sc["sh_degree"], bg)
    for gi, g in enumerate(gazes):
        rep = {"variant": "smfr", "tag": tag, "P": P, "W": W, "H": H, "gaze": list(g)}
        try:
            if mod is None:
                raise RuntimeError("reference module ref_naive_C is not built")
            gaze = torch.tensor(g, dtype=torch.float32, device="cuda")
            n_r, col_r, rad_r, geom, binning, img = ref_api.smfr_forward(mod, sc, c, gaze)
            torch.cuda.synchronize()
            br = ref_api.decode_binning(binning, n_r)
            ir = ref_api.decode_img(img, W, H)
            n_o, col_o, rad_o, pl_o, rg_o, item = ops.forward_smfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                                   sc["highest_levels"], gaze, 0.05, True, rs, want_lists=True)
            torch.cuda.synchronize()
            rep["num_rendered_ref"] = int(n_r); rep["num_rendered_ours"] = int(n_o)
            rep["radii_mismatch"] = int((rad_r != rad_o).sum().item())
            rep["point_list_mismatch"] = int((br["point_list"][:n_r] != pl_o[:n_r]).sum().item()) if n_r == n_o else -1
            T = ((W + 15) // 16) * ((H + 15) // 16)
            rep["ranges_mismatch"] = int((ir["ranges"][:T] != rg_o[:T]).sum().item())
            d = (col_r - col_o).abs()
            rep["img_max_abs"] = float(d.max().item()); rep["img_n_gt_1e-6"] = int((d > 1e-6).sum().item())
            _, col_l, _ = ops.forward_smfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"], sc["highest_levels"],
                                           gaze, 0.05, True, rs)
            rep["lazy_img_max_abs"] = float((col_l - col_r).abs().max().item()); rep["lazy_stats"] = dict(ops.last_stats)
            if golden_dir and gi < 2:
                np.savez_compressed(os.path.join(golden_dir, f"smfr_{tag}_g{gi}.npz"), gaze=np.array(g, np.float32),
                                    color=col_r.cpu().numpy(), radii=rad_r.cpu().numpy(), num_rendered=np.int64(n_r),
                                    point_list=br["point_list"].cpu().numpy(), ranges=ir["ranges"][:T].cpu().numpy())
            if do_time and gi == 0:
                rep["time_ref"] = time_fn(lambda: ref_api.smfr_forward(mod, sc, c, gaze))
                rep["time_ours"] = time_fn(lambda: ops.forward_smfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                                    sc["highest_levels"], gaze, 0.05, True, rs))
        except Exception as ex:
            rep["error"] = repr(ex); rep["trace"] = traceback.format_exc()[-1500:]
        rep_list.append(rep)
        print(json.dumps({k: v for k, v in rep.items() if k != "trace"}), flush=True)


def run_mmfr(scn, cam, gazes, rep_list, do_time, golden_dir, tag):
    """MMFR baseline (mmfr_pcheck_obb) vs ref_mmfr_C: four level calls per gaze (level 0 first: the reference refreshes
    its static tile tables only then); the level models are nested subsets of the scene (every 1, 2, 4, 8-th Gaussian)."""
    mod = ref_api.ref_module("ref_mmfr_C")
    full = to_cuda(scn)
    c = to_cuda(cam)
    W, H = cam["image_width"], cam["image_height"]
    bg = torch.zeros(3, device="cuda")
    rs = settings(c, full["sh_degree"], bg)
    levels = []
    for l in range(4):
        sub = {k: (v[:: 1 << l].contiguous() if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == full["means3D"].shape[0] else v)
               for k, v in full.items()}
        levels.append(sub)
    for gi, g in enumerate(gazes):
        gaze = torch.tensor(g, dtype=torch.float32, device="cuda")
        for l, sc in enumerate(levels):
            P = sc["means3D"].shape[0]
            rep = {"variant": "mmfr", "tag": tag, "P": P, "W": W, "H": H, "gaze": list(g), "cur_level": l}
            try:
                if mod is None:
                    raise RuntimeError("reference module ref_mmfr_C is not built")
                n_r, col_r, rad_r, geom, binning, img = ref_api.mmfr_forward(mod, sc, c, l, gaze)
                torch.cuda.synchronize()
                br = ref_api.decode_binning(binning, n_r)
                ir = ref_api.decode_img(img, W, H)
                n_o, col_o, rad_o, pl_o, rg_o, item = ops.forward_mmfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                                       l, gaze, 0.05, True, rs, want_lists=True)
                torch.cuda.synchronize()
                rep["num_rendered_ref"] = int(n_r); rep["num_rendered_ours"] = int(n_o)
                rep["radii_mismatch"] = int((rad_r != rad_o).sum().item())
                rep["point_list_mismatch"] = int((br["point_list"][:n_r] != pl_o[:n_r]).sum().item()) if n_r == n_o else -1
                T = ((W + 15) // 16) * ((H + 15) // 16)
                rep["ranges_mismatch"] = int((ir["ranges"][:T] != rg_o[:T]).sum().item())
                d = (col_r - col_o).abs()
                rep["img_max_abs"] = float(d.max().item()); rep["img_n_gt_1e-6"] = int((d > 1e-6).sum().item())
                _, col_l, _ = ops.forward_mmfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"], l, gaze, 0.05, True, rs)
                rep["lazy_img_max_abs"] = float((col_l - col_r).abs().max().item())
                if golden_dir and gi < 1:
                    np.savez_compressed(os.path.join(golden_dir, f"mmfr_{tag}_g{gi}_l{l}.npz"), gaze=np.array(g, np.float32), cur_level=np.int64(l),
                                        color=col_r.cpu().numpy(), radii=rad_r.cpu().numpy(), num_rendered=np.int64(n_r),
                                        point_list=br["point_list"].cpu().numpy(), ranges=ir["ranges"][:T].cpu().numpy())
                if do_time and gi == 0:
                    rep["time_ref"] = time_fn(lambda: ref_api.mmfr_forward(mod, sc, c, l, gaze))
                    rep["time_ours"] = time_fn(lambda: ops.forward_mmfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"],
                                                                        l, gaze, 0.05, True, rs))
            except Exception as ex:
                rep["error"] = repr(ex); rep["trace"] = traceback.format_exc()[-1500:]
            rep_list.append(rep)
            print(json.dumps({k: v for k, v in rep.items() if k != "trace"}), flush=True)


PS1_VARIANTS = {"obb": ("ref_obb_C", ops.MODE_OBB), "sum": ("ref_sum_C", ops.MODE_SUM), "max": ("ref_max_C", ops.MODE_MAX),
                "lwmc": ("ref_lwmc_C", ops.MODE_LWMC), "vanilla": ("ref_vanilla_C", ops.MODE_VANILLA)}


def run_ps1(variant, scn, cam, rep_list, do_time, golden_dir, tag):
    sum_mode = variant != "obb"          # the training family: sum / max / lwmc / vanilla (forward state + backward)
    stat_mode = sum_mode and variant != "vanilla"   # ... of which all but the stock rasterizer return per-Gaussian statistics
    mod = ref_api.ref_module(PS1_VARIANTS[variant][0])
    sc = to_cuda(scn)
    c = to_cuda(cam)
    W, H = cam["image_width"], cam["image_height"]
    P = sc["means3D"].shape[0]
    bg = torch.zeros(3, device="cuda")
    rs = settings(c, sc["sh_degree"], bg)
    mode = PS1_VARIANTS[variant][1]
    rep = {"variant": variant, "tag": tag, "P": P, "W": W, "H": H}
    loss_map = None
    if variant == "lwmc":
        loss_map = torch.from_numpy(np.random.default_rng(17).random((H, W)).astype(np.float32)).cuda()
    try:
        if mod is None:
            raise RuntimeError(f"reference module {PS1_VARIANTS[variant][0]} is not built")
        res = ref_api.ps1_forward(mod, sc, c, loss_map=loss_map)
        torch.cuda.synchronize()
        n_r, col_r, rad_r, geom, binning, img = res[:6]
        gr = ref_api.decode_geom(geom, P, "ps1")
        br = ref_api.decode_binning(binning, n_r)
        ir = ref_api.decode_img(img, W, H)
        ref = {"n": n_r, "color": col_r, "radii": rad_r, "point_list": br["point_list"], "ranges": ir["ranges"]}
        out = ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs, want_lists=True,
                              loss_map=loss_map)
        torch.cuda.synchronize()
        n_o, col_o, rad_o, item = out[:4]
        pl_o, rg_o = out[-2], out[-1]
        go = ops.geometry(item, mode, P, W, H)
        ours = {"n": n_o, "color": col_o, "radii": rad_o, "point_list": pl_o, "ranges": rg_o}
        compare_common(rep, ref, ours, gr, go, W, H)
        if not sum_mode:
            lz = ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs)
            rep["lazy_img_max_abs"] = float((lz[1] - col_r).abs().max().item()); rep["lazy_stats"] = dict(ops.last_stats)
        vis = (rad_r > 0)
        rep["rgb_max_abs"] = float(((gr["rgb"] - go["rgb"]).abs() * vis.unsqueeze(-1)).max().item())
        if sum_mode:
            rep["cov3D_bit_mismatch"] = mism(gr["cov3D"], go["cov3D"], vis)
            rep["n_contrib_mismatch"] = None
            # the lazy (consumption-driven) training kernel: same statistics, same image, same saved state
            lz = ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs,
                                 loss_map=loss_map)
            torch.cuda.synchronize()
            rep["lazy_img_max_abs"] = float((lz[1] - col_r).abs().max().item())
            if stat_mode:
                gc_r, ct_r = res[6], res[7]
                gc_o, ct_o = out[4], out[5]
                rep["gaussians_count_mismatch"] = int((gc_r != gc_o).sum().item())
                rep["contrib_max_rel"] = float(((ct_r - ct_o).abs() / (ct_r.abs() + 1e-6)).max().item())
                rep["lazy_gaussians_count_mismatch"] = int((gc_r != lz[4]).sum().item())
                rep["lazy_contrib_max_rel"] = float(((ct_r - lz[5]).abs() / (ct_r.abs() + 1e-6)).max().item())
                rep["lazy_contrib_bit_mismatch"] = mism(ct_r, lz[5])
                rep["contrib_bit_mismatch"] = mism(ct_r, ct_o)
            # backward
            grad_out = torch.from_numpy(np.random.default_rng(3).standard_normal((3, H, W)).astype(np.float32)).cuda()
            g_ref = ref_api.ps1_backward(mod, sc, c, rad_r, grad_out, geom, n_r, binning, img)
            g_our = ops.backward_ps1(item, sc["means3D"], rad_o, sc["scales"], sc["rotations"], None, sc["shs"], None, rs, grad_out)
            g_lazy = ops.backward_ps1(lz[3], sc["means3D"], lz[2], sc["scales"], sc["rotations"], None, sc["shs"], None, rs, grad_out)
            torch.cuda.synchronize()
            names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
            gcmp = {}
            for nm, a, b, l in zip(names, g_ref, g_our, g_lazy):
                num = (a - b).norm().item(); den = a.norm().item()
                gcmp[nm] = {"rel_l2": num / (den + 1e-30), "max_abs": float((a - b).abs().max().item()), "ref_max": float(a.abs().max().item()),
                            "lazy_rel_l2": (a - l).norm().item() / (den + 1e-30)}
            rep["grads"] = gcmp
            if do_time:
                rep["time_bwd_ref"] = time_fn(lambda: ref_api.ps1_backward(mod, sc, c, rad_r, grad_out, geom, n_r, binning, img), 3, 10)
                rep["time_bwd_ours"] = time_fn(lambda: ops.backward_ps1(item, sc["means3D"], rad_o, sc["scales"], sc["rotations"], None, sc["shs"], None, rs, grad_out), 3, 10)
            if golden_dir:
                np.savez_compressed(os.path.join(golden_dir, f"{variant}_{tag}_bwd.npz"), grad_seed=np.int64(3), numpy_grad=np.int64(1),
                                    **{nm: a.cpu().numpy() for nm, a in zip(names, g_ref)})
        rep["stats"] = dict(ops.last_stats)
        if golden_dir:
            extra = {}
            if sum_mode:
                extra = {"final_T": ir["accum_alpha"].cpu().numpy(), "n_contrib": ir["n_contrib"].cpu().numpy()}
                if stat_mode:
                    extra.update({"gaussians_count": res[6].cpu().numpy(), "contributions": res[7].cpu().numpy()})
                if loss_map is not None:
                    extra["loss_map_seed"] = np.int64(17)
            np.savez_compressed(os.path.join(golden_dir, f"{variant}_{tag}.npz"), color=col_r.cpu().numpy(), radii=rad_r.cpu().numpy(),
                                num_rendered=np.int64(n_r), point_list=br["point_list"].cpu().numpy(),
                                ranges=ir["ranges"][: ((W + 15) // 16) * ((H + 15) // 16)].cpu().numpy(),
                                means2D=gr["means2D"].cpu().numpy(), depths=gr["depths"].cpu().numpy(), conic=gr["conic"].contiguous().cpu().numpy(), **extra)
        if do_time:
            rep["time_ref"] = time_fn(lambda: ref_api.ps1_forward(mod, sc, c, loss_map=loss_map))
            rep["time_ours"] = time_fn(lambda: ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs,
                                                               loss_map=loss_map))
    except Exception as ex:
        rep["error"] = repr(ex); rep["trace"] = traceback.format_exc()[-1500:]
    rep_list.append(rep)
    print(json.dumps({k: v for k, v in rep.items() if k != "trace"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="small")
    ap.add_argument("--variants", default="fov,obb,sum")
    ap.add_argument("--golden", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--gazes", type=int, default=9)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_report.json"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    golden_dir = None
    if a.golden:
        golden_dir = os.path.join(ROOT, "gpurun_out", "golden")
        os.makedirs(golden_dir, exist_ok=True)
    reports = []
    for size in a.sizes.split(","):
        t0 = time.time()
        if size == "small":
            scn = synth.make_scene_cube(10000, 0); cams = [synth.config1_camera()]
        elif size == "mid":
            scn = synth.make_scene_bicycle(300000, 1, log_scale_mu=-3.6); cams = [synth.ring_cameras(30, 800, 600)[3]]
        elif size == "big":
            scn = synth.make_scene_bicycle(6000000, 1); cams = [synth.ring_cameras(30)[0]]
        else:
            raise SystemExit("unknown size " + size)
        print(f"# scene {size} generated in {time.time() - t0:.1f}s", flush=True)
        for ci, cam in enumerate(cams):
            tag = f"{size}_c{ci}"
            gd = golden_dir if size == "small" else None
            for v in a.variants.split(","):
                if v == "fov":
                    run_fov(scn, cam, synth.GAZES_9[: a.gazes], reports, a.time, gd, tag)
                elif v == "smfr":
                    run_smfr(scn, cam, synth.GAZES_9[: a.gazes], reports, a.time, gd, tag)
                elif v == "mmfr":
                    run_mmfr(scn, cam, synth.GAZES_9[: a.gazes], reports, a.time, gd, tag)
                else:
                    run_ps1(v, scn, cam, reports, a.time, gd, tag)
        json.dump(reports, open(a.out, "w"), indent=1)
    print("# report written to", a.out)


if __name__ == "__main__":
    main()
