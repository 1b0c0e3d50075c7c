#!/usr/bin/env python
"""Where does the CPU oracle's (Gaussian, tile) instance set differ from the CUDA library's at full size?

VERDICT r1: frame 0 of the bench workload (camera 0, gaze (0.25, 0.25)) has 8 296 084 instances in the reference binary and
in our CUDA path, 8 296 083 in the CPU oracle.  This tool finds the differing instances and dumps everything needed to
reproduce the decision on a CPU-only machine (the Gaussian's inputs, the camera, the tile, both sides' projections and the
tile tables at that tile) into gpurun_out/oracle_diff/frame<k>.npz + .json.

  python tools/oracle_diff.py [--frames 0,1,2] [--size big]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bench  # noqa: E402
from fovgs import ops  # noqa: E402
import oracle  # noqa: E402
import parity_gpu  # noqa: E402


def instance_keys(point_list, ranges):
    rg = np.asarray(ranges, np.int64)
    n = rg[:, 1] - rg[:, 0]
    tile = np.repeat(np.arange(rg.shape[0], dtype=np.int64), np.maximum(n, 0))
    return (tile << 32) | np.asarray(point_list, np.int64)[: tile.size]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", default="0,1,2")
    ap.add_argument("--size", default="big")
    a = ap.parse_args()
    out_dir = os.path.join(ROOT, "gpurun_out", "oracle_diff")
    os.makedirs(out_dir, exist_ok=True)
    wl = bench.Workload(a.size)
    sc = parity_gpu.to_cuda(wl.scene)
    bg = torch.zeros(3, device="cuda")
    summary = []
    for f in [int(x) for x in a.frames.split(",")]:
        cam, gaze = wl.frame(f)
        c = parity_gpu.to_cuda(cam)
        rs = parity_gpu.settings(c, wl.scene["sh_degree"], bg)
        g = torch.tensor(gaze, dtype=torch.float32, device="cuda")
        n, color, radii, pl, rg, item = ops.forward_fov(sc["means3D"], sc["opacities4"], sc["scales"], sc["rotations"], sc["shs_rest"],
                                                        sc["shs_dcs"], sc["highest_levels"], g, 0.05, True, rs, want_lists=True)
        torch.cuda.synchronize()
        lvl, mn, gx, gy, bl = ops.fov_tile_tables(item, wl.W, wl.H)
        geo = ops.geometry(item, ops.MODE_FOV, wl.P, wl.W, wl.H)
        o = oracle.forward_fov(wl.scene, cam, gaze, list_cap=1 << 27)
        ko = instance_keys(o["point_list"], o["ranges"])
        kg = instance_keys(pl.cpu().numpy(), rg.cpu().numpy())
        only_gpu = np.setdiff1d(kg, ko)
        only_cpu = np.setdiff1d(ko, kg)
        ids = np.unique(np.concatenate([only_gpu, only_cpu]) & 0xffffffff).astype(np.int64)
        rec = {"frame": f, "gaze": list(gaze), "n_gpu": int(n), "n_oracle": int(o["num_rendered"]),
               "only_gpu": [[int(k >> 32), int(k & 0xffffffff)] for k in only_gpu[:64]],
               "only_oracle": [[int(k >> 32), int(k & 0xffffffff)] for k in only_cpu[:64]],
               "radii_mismatch": int((radii.cpu().numpy() != o["radii"]).sum()),
               "tile_min_bit_mismatch": int((mn.cpu().numpy().view(np.int32) != o["tile_min"].view(np.int32)).sum()),
               "tile_level_bit_mismatch": int((lvl.cpu().numpy().view(np.int32) != o["tile_level"].view(np.int32)).sum()),
               "img_max_abs": float(np.abs(color.cpu().numpy() - o["color"]).max())}
        print(json.dumps(rec), flush=True)
        summary.append(rec)
        idt = torch.from_numpy(ids).cuda()
        np.savez_compressed(
            os.path.join(out_dir, f"frame{f}.npz"), ids=ids, gaze=np.asarray(gaze, np.float32),
            only_gpu=only_gpu, only_oracle=only_cpu,
            **{k: np.asarray(v) for k, v in cam.items() if isinstance(v, (np.ndarray, int, float))},
            **{"in_" + k: wl.scene[k][ids] for k in ("means3D", "scales", "rotations", "highest_levels", "opacities4", "shs_dcs", "shs_rest")},
            gpu_means2D=geo["means2D"][idt].cpu().numpy(), gpu_conic=geo["conic"][idt].cpu().numpy(), gpu_depths=geo["depths"][idt].cpu().numpy(),
            gpu_radii=radii[idt].cpu().numpy(), orc_means2D=o["means2D"][ids], orc_conic=o["conic"][ids], orc_depths=o["depths"][ids],
            orc_radii=o["radii"][ids], gpu_tile_min=mn.cpu().numpy(), orc_tile_min=o["tile_min"], gpu_tile_level=lvl.cpu().numpy(),
            orc_tile_level=o["tile_level"])
    json.dump(summary, open(os.path.join(out_dir, "summary.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
