#!/usr/bin/env python
"""Per-tile CTA timeline of k_lazy_blend (needs a library built with `make EXTRA=-DFOVGS_TILE_TIMING`):
makespan vs per-SM busy time, time share by tile size class, longest tiles."""
import ctypes as C, os, sys, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from fovgs import ops, synth, _lib
from parity_gpu import to_cuda, settings

def main():
    scn = synth.make_scene_bicycle(6000000, 1); cams = synth.ring_cameras(30)
    sc = to_cuda(synth.add_foveation(scn)); bg = torch.zeros(3, device="cuda")
    L = _lib.lib()
    T = 120 * 68
    for f in (0, 4, 13):
        c = to_cuda(cams[f % 30]); rs = settings(c, sc["sh_degree"], bg)
        gaze = torch.tensor(synth.GAZES_9[f % 9], dtype=torch.float32, device="cuda")
        for _ in range(3):
            ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"], sc["shs_dcs"], sc["highest_levels"], gaze, 0.05, True, rs)
        torch.cuda.synchronize()
        buf = np.zeros(4 * T, np.uint32)
        assert L.fovgs_debug_tile_times(buf.ctypes.data_as(C.c_void_p), T) == 0
        a = buf.reshape(T, 4).astype(np.int64)
        t0 = a[:, 0].min(); dur = ((a[:, 1] - a[:, 0]) & 0xffffffff) / 1e3; start = (a[:, 0] - t0) / 1e3; end = start + dur
        n = a[:, 3]
        print(f"frame {f} gaze {synth.GAZES_9[f % 9]}: makespan {end.max():.1f} us, sum tile-us {dur.sum():.0f}, mean concurrency {dur.sum() / end.max():.0f} of {148 * 4}")
        for lo, hi in [(0, 1), (1, 256), (256, 2048), (2048, 8192), (8192, 1 << 30)]:
            m = (n >= lo) & (n < hi)
            if m.any():
                print(f"   n in [{lo},{hi}): tiles {m.sum()}, time share {100 * dur[m].sum() / dur.sum():.1f}%, mean {dur[m].mean():.1f} us, max {dur[m].max():.1f} us")
        busy = collections.defaultdict(float); fin = collections.defaultdict(float)
        for s_, d, e in zip(a[:, 2], dur, end):
            busy[s_] += d; fin[s_] = max(fin[s_], e)
        b = np.array(list(busy.values())); fi = np.array(list(fin.values()))
        print(f"   per-SM busy tile-us min/mean/max {b.min():.0f}/{b.mean():.0f}/{b.max():.0f}; per-SM finish us min/mean/max {fi.min():.0f}/{fi.mean():.0f}/{fi.max():.0f}")
        edges = np.linspace(0, end.max(), 11)
        print("   alive tiles per tenth:", [int(((start < edges[i + 1]) & (end > edges[i])).sum()) for i in range(10)])
        o = np.argsort(-dur)[:6]
        print("   longest:", [(int(n[i]), round(float(dur[i]), 1), round(float(start[i]), 1)) for i in o])

if __name__ == "__main__":
    main()
