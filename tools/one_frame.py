#!/usr/bin/env python
"""Render a few frames of one config with our library only (for ncu launch lists / single-kernel captures).
   python tools/one_frame.py --size big --variant fov --frames 3"""
import argparse, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
from fovgs import ops, synth  # noqa: E402
ops.set_full_stats(True)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from parity_gpu import to_cuda, settings  # noqa: E402

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="mid"); ap.add_argument("--variant", default="fov")
    ap.add_argument("--frames", type=int, default=3); ap.add_argument("--ref", action="store_true")
    a = ap.parse_args()
    if a.size == "small": scn = synth.make_scene_cube(10000, 0); cam = synth.config1_camera()
    elif a.size == "mid": scn = synth.make_scene_bicycle(300000, 1, log_scale_mu=-3.6); cam = synth.ring_cameras(30, 800, 600)[3]
    else: scn = synth.make_scene_bicycle(6000000, 1); cam = synth.ring_cameras(30)[0]
    c = to_cuda(cam); bg = torch.zeros(3, device="cuda")
    if a.variant == "fov":
        sc = to_cuda(synth.add_foveation(scn)); rs = settings(c, sc["sh_degree"], bg)
        gaze = torch.tensor([0.5, 0.5], device="cuda")
        if a.ref:
            sys.path.insert(0, os.path.join(ROOT, "oracle")); import ref_api
            mod = ref_api.ref_module("ref_fov_C"); fn = lambda: ref_api.fov_forward(mod, sc, c, gaze)
        else:
            fn = lambda: ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"], sc["highest_levels"], gaze, 0.05, True, rs)
    else:
        sc = to_cuda(scn); rs = settings(c, sc["sh_degree"], bg)
        mode = ops.MODE_SUM if a.variant == "sum" else ops.MODE_OBB
        if a.ref:
            sys.path.insert(0, os.path.join(ROOT, "oracle")); import ref_api
            mod = ref_api.ref_module("ref_sum_C" if a.variant == "sum" else "ref_obb_C"); fn = lambda: ref_api.ps1_forward(mod, sc, c)
        else:
            fn = lambda: ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs)
    for _ in range(a.frames):
        fn(); torch.cuda.synchronize()
    print("stats", ops.last_stats)
if __name__ == "__main__":
    main()
