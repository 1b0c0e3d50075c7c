#!/usr/bin/env python
"""Frames/s under the REFERENCE'S OWN timing protocol (render_compose_gazes_fps.py:25-73, gaussian_renderer_fov/__init__.py:
74-97): for each of the 9 gazes, 10 warm-up renders of view 0, then for every view 5 renders, each bracketed by a CUDA event
pair around the rasterizer call alone and followed by torch.cuda.synchronize(); fps_view = 5 / (sum of the 5 times);
mean over views, then mean over gazes.  Synthetic bench workload (6 M Gaussians, 1920x1080, 30 ring cameras).

  python tools/fps_protocol.py [--views 30] [--impl ours|reference|both] [--size big|ref]
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402


def protocol(render, views, gazes):
    starter, ender = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_gaze = []
    for g in gazes:
        for _ in range(10):
            render(views[0], g, None, None)
            torch.cuda.synchronize()
        fpss = []
        for v in views:
            t = 0.0
            for _ in range(5):
                render(v, g, starter, ender)
                torch.cuda.synchronize()
                t += starter.elapsed_time(ender)
            fpss.append(5 / (t / 1000))
        per_gaze.append(sum(fpss) / len(fpss))
    return per_gaze


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=30)
    ap.add_argument("--impl", default="both")
    ap.add_argument("--size", default="big", help="big = bench workload; ref = 1.16 M Gaussians at 1237x822 (the published 702 FPS case)")
    ap.add_argument("--base_folder", help="a trained model instead of the synthetic workload: the folder render_compose_gazes_fps.py "
                    "takes (1_PS1_<L>_<S>/point_cloud/iteration_55000/point_cloud.ply + composed_<L>_<S>/*.pt)")
    ap.add_argument("--layer_num", type=int, default=4); ap.add_argument("--max_pooling_size", type=int, default=4)
    ap.add_argument("--cameras", help="cameras.json of the model (with --base_folder); the test views come first in it "
                    "(scene/__init__.py:55-61), so --views 25 takes bicycle's 25 test views")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    if a.base_folder:
        from fovgs import io, synth
        scene = io.load_foveated_model(a.base_folder, a.layer_num, a.max_pooling_size)
        cams_np = io.cameras_from_json(a.cameras)[: a.views]
        name, gz = f"{os.path.basename(os.path.normpath(a.base_folder))}_{scene['means3D'].shape[0]}g", synth.GAZES_9
    else:
        wl = bench.Workload(a.size)
        scene, cams_np, name, gz = wl.scene, wl.cams[: a.views], wl.name, wl.gazes
    sc = bench.to_dev(scene, dev); bg = torch.zeros(3, device=dev)
    cams = [bench.to_dev(c, dev) for c in cams_np]
    gazes = [torch.tensor([g[0], g[1]]).float().cuda() for g in gz]
    out = {"workload": name}
    if a.impl in ("ours", "both"):
        import diff_gaussian_rasterization_fov_pcheck_obb as pkg

        def render(c, g, s, e):
            rs = pkg.GaussianRasterizationSettings(c["image_height"], c["image_width"], c["tanfovx"], c["tanfovy"], bg, 1.0,
                                                   c["viewmatrix"], c["projmatrix"], 3, c["campos"], False, False)
            r = pkg.GaussianRasterizer(raster_settings=rs)
            if s is not None: s.record()
            img, radii = r(means3D=sc["means3D"], means2D=None, opacities=sc["opacities4"], shs_rest=sc["shs_rest"], scales=sc["scales"],
                           rotations=sc["rotations"], shs_dcs=sc["shs_dcs"], highest_levels=sc["highest_levels"], gazeArray=g, alpha=0.05,
                           blending=True)
            if e is not None: e.record()
            return img

        with torch.no_grad():
            pg = protocol(render, cams, gazes)
        out["ours"] = {"fps": float(np.mean(pg)), "per_gaze": [round(x, 1) for x in pg]}
    if a.impl in ("reference", "both"):
        import ref_api
        mod = ref_api.ref_module("ref_fov_C")
        if mod is not None:
            def render_ref(c, g, s, e):
                if s is not None: s.record()
                o = ref_api.fov_forward(mod, sc, c, g, 0.05, True, bg)
                if e is not None: e.record()
                return o[1]
            with torch.no_grad():
                pg = protocol(render_ref, cams, gazes)
            out["reference"] = {"fps": float(np.mean(pg)), "per_gaze": [round(x, 1) for x in pg]}
    out["protocol"] = "render_compose_gazes_fps.py: 9 gazes x %d views, 10 warm-ups per gaze, 5 event-timed renders per view, sync after each" % len(cams)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
