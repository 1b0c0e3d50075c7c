#!/usr/bin/env python
"""Where does the end-to-end loop of bench.py spend its time?  Variants: read-back depth, no read-back, device-resident camera."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
import bench
from fovgs import ops
import diff_gaussian_rasterization_fov_pcheck_obb as fovpkg

def main():
    dev = torch.device("cuda", 0)
    wl = bench.Workload("big")
    sc = bench.to_dev(wl.scene, dev); bg = torch.zeros(3, device=dev)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    cams_host = [{k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in c.items()} for c in wl.cams]
    gazes_host = [pin(np.asarray(g, np.float32)) for g in wl.gazes]
    def settings(c):
        return fovpkg.GaussianRasterizationSettings(c["image_height"], c["image_width"], c["tanfovx"], c["tanfovy"], bg, 1.0,
                                                    c["viewmatrix"], c["projmatrix"], wl.scene["sh_degree"], c["campos"], False, False)
    def render(rs, gaze):
        r = fovpkg.GaussianRasterizer(raster_settings=rs)
        return r(means3D=sc["means3D"], means2D=None, opacities=sc["opacities4"], shs_rest=sc["shs_rest"], scales=sc["scales"],
                 rotations=sc["rotations"], shs_dcs=sc["shs_dcs"], highest_levels=sc["highest_levels"], gazeArray=gaze, alpha=0.05, blending=True)
    cams_dev = [bench.to_dev(c, dev) for c in wl.cams]; gz_dev = [torch.tensor(g, dtype=torch.float32, device=dev) for g in wl.gazes]
    N = 180
    for deferred in (True, False):
        ops.set_deferred_check(deferred)
        for name, depth, h2d in (("depth2", 2, True), ("depth4", 4, True), ("depth8", 8, True), ("no_readback", 0, True), ("dev_camera_depth4", 4, False)):
            rb = bench.Readback(dev, (3, wl.H, wl.W), depth=depth) if depth else None
            cpu = 0.0
            with torch.no_grad():
                for phase in range(2):
                    torch.cuda.synchronize(); t0 = time.perf_counter(); cpu = 0.0
                    for f in range(N):
                        c0 = time.perf_counter()
                        if h2d:
                            c = cams_host[f % 30]
                            cd = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
                            g = gazes_host[f % 9].to(dev, non_blocking=True)
                        else:
                            cd = cams_dev[f % 30]; g = gz_dev[f % 9]
                        img, _ = render(settings(cd), g)
                        c1 = time.perf_counter()
                        if rb: rb.push(img)
                        cpu += c1 - c0
                    if rb: rb.drain()
                    torch.cuda.synchronize(); dt = time.perf_counter() - t0
            ops.check_pending(dev) if deferred else None
            print(f"deferred={deferred} {name}: {N / dt:.1f} frames/s, cpu per frame before push {1e6 * cpu / N:.0f} us", flush=True)
    ops.set_deferred_check(False)

if __name__ == "__main__":
    main()
