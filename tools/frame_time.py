#!/usr/bin/env python
"""Back-to-back frame time of the foveated bench workload WITHOUT stage events (the library's profile mode records an event
between every two kernels, which also breaks the chain of programmatic dependent launches), then the stage table of the same
frames with the events on.  For A/B runs of alternative builds: FOVGS_LIB_PATH=... python tools/frame_time.py [--no-pdl]
   python tools/frame_time.py --frames 60 --variant fov|obb|sum"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
from fovgs import ops, synth  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from parity_gpu import to_cuda, settings  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60); ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--variant", default="fov"); ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--gaussians", type=int, default=6000000)
    a = ap.parse_args()
    if a.no_pdl:
        from fovgs._lib import lib
        assert lib().fovgs_set_option(3, 1) == 0
    scn = synth.make_scene_bicycle(a.gaussians, 1)
    cams = [to_cuda(c) for c in synth.ring_cameras(30)]
    bg = torch.zeros(3, device="cuda")
    sc = to_cuda(synth.add_foveation(scn)) if a.variant == "fov" else to_cuda(scn)
    rs = [settings(c, sc["sh_degree"], bg) for c in cams]
    gazes = [torch.tensor(g, dtype=torch.float32, device="cuda") for g in synth.GAZES_9]

    def frame(f):
        if a.variant == "fov":
            return ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"],
                                   sc["highest_levels"], gazes[f % 9], 0.05, True, rs[f % 30])
        mode = ops.MODE_SUM if a.variant == "sum" else ops.MODE_OBB
        return ops.forward_ps1(mode, sc["means3D"], sc["opacity"], sc["scales"], sc["rotations"], None, sc["shs"], None, rs[f % 30])

    out = {"variant": a.variant, "frames": a.frames, "pdl": not a.no_pdl, "lib": os.environ.get("FOVGS_LIB_PATH", "default")}
    with torch.no_grad():
        for f in range(a.warmup):
            frame(f)
        torch.cuda.synchronize()
        ops.set_deferred_check(True)
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for f in range(a.warmup, a.warmup + a.frames):
                frame(f)
            e1.record(); torch.cuda.synchronize()
            out["ms_per_frame_run%d" % rep] = e0.elapsed_time(e1) / a.frames
        ops.check_pending(torch.device("cuda"))
        ops.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in range(a.warmup, a.warmup + a.frames):
            frame(f)
        e1.record(); torch.cuda.synchronize()
        out["ms_per_frame_with_stage_events"] = e0.elapsed_time(e1) / a.frames
        st = ops.profile_read_all()[-a.frames:]
        out["stages"] = {k: round(float(np.mean([s[k] for s in st])), 4) for k in ops.STAGE_NAMES}
        ops.check_pending(torch.device("cuda"))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
