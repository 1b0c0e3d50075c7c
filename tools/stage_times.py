#!/usr/bin/env python
"""Mean stage times (library CUDA events) + whole-call times for one variant on the big config.
   python tools/stage_times.py --variant sum --frames 8 [--backward]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
from fovgs import ops, synth  # noqa: E402
ops.set_full_stats(True)   # these tools print the blend stage's counters: wait for the end of each frame
sys.path.insert(0, os.path.join(ROOT, "tools"))
from parity_gpu import to_cuda, settings  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="sum"); ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--size", default="big"); ap.add_argument("--backward", action="store_true")
    ap.add_argument("--split", default="0", help="comma list of FOVGS_OPT_SPLIT_STAGES settings to run, e.g. 0,1")
    ap.add_argument("--first", type=int, default=0, help="index of the first measured frame (bench.py frame numbering: camera f%30, gaze f%9)")
    ap.add_argument("--no-pdl", action="store_true", help="FOVGS_OPT_NO_PDL: the two blend launches run back to back")
    a = ap.parse_args()
    if a.no_pdl:
        from fovgs._lib import lib
        assert lib().fovgs_set_option(3, 1) == 0
    if a.size == "big": scn = synth.make_scene_bicycle(6000000, 1); cams = synth.ring_cameras(30)
    else: scn = synth.make_scene_bicycle(300000, 1, log_scale_mu=-3.6); cams = synth.ring_cameras(30, 800, 600)
    bg = torch.zeros(3, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    if a.variant in ("fov", "smfr"):
        sc = to_cuda(synth.add_foveation(scn))
    else:
        sc = to_cuda(scn)
    ops.profile_enable(True)
    run(a, sc, cams, bg, ev, 0)


def run(a, sc, cams, bg, ev, split):
    fwd_ms, bwd_ms = [], []
    for i in range(a.frames + 2):
        f = max(a.first - 2, 0) + i if a.first >= 2 else i
        c = to_cuda(cams[f % 30]); rs = settings(c, sc["sh_degree"], bg)
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        if a.variant == "fov":
            gaze = torch.tensor(synth.GAZES_9[f % 9], dtype=torch.float32, device="cuda")
            out = ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"], sc["highest_levels"], gaze, 0.05, True, rs)
        elif a.variant == "smfr":
            gaze = torch.tensor(synth.GAZES_9[f % 9], dtype=torch.float32, device="cuda")
            out = ops.forward_smfr(sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], sc["shs"], sc["highest_levels"], gaze, 0.05, True, rs)
        else:
            mode = ops.MODE_SUM if a.variant == "sum" else ops.MODE_OBB
            out = ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs)
        e1.record()
        if a.backward and a.variant == "sum":
            g = torch.from_numpy(np.random.default_rng(3).standard_normal(out[1].shape).astype(np.float32)).cuda()
            e1.record()
            ops.backward_ps1(out[3], sc["means3D"], out[2], sc["scales"], sc["rotations"], None, sc["shs"], None, rs, g)
        e2.record(); torch.cuda.synchronize()
        if i >= 2: fwd_ms.append(e0.elapsed_time(e1)); bwd_ms.append(e1.elapsed_time(e2))
    st = ops.profile_read_all()[2:]
    mean = {k: float(np.mean([s[k] for s in st])) for k in ops.STAGE_NAMES}
    print("split", split, "variant", a.variant, "fwd_ms_mean", np.mean(fwd_ms), "bwd_ms_mean", np.mean(bwd_ms), "stages", {k: round(v, 4) for k, v in mean.items()}, "stats", ops.last_stats)


if __name__ == "__main__":
    main()
