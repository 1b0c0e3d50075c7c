#!/usr/bin/env python
"""Golden camera matrices from the REFERENCE's own functions (run in the build container, where /root/reference exists):
utils/graphics_utils.py getWorld2View2 / getProjectionMatrix / focal2fov / fov2focal, combined as scene/cameras.py:54-57 and
utils/camera_utils.py:62-82 (camera_to_JSON) combine them.  Output: tests/golden/cameras_ref.npz (+ the JSON entries)."""
import importlib.util
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FOVGS_REFERENCE_ROOT", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_graphics_utils", os.path.join(REF, "fov3dgs", "utils", "graphics_utils.py"))
gu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gu)


def main():
    rng = np.random.default_rng(123)
    out = {}
    entries = []
    for i in range(4):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        T = rng.normal(size=3) * 3.0
        W, H = [(1920, 1080), (1237, 822), (640, 360), (256, 256)][i]
        fx = W / (2 * math.tan(math.radians(30 + 5 * i)))
        fovx = gu.focal2fov(fx, W); fovy = gu.focal2fov(fx, H)
        wv = torch.tensor(gu.getWorld2View2(R, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
        pj = gu.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)
        full = (wv.unsqueeze(0).bmm(pj.unsqueeze(0))).squeeze(0)
        center = wv.inverse()[3, :3]
        # camera_to_JSON (utils/camera_utils.py:62-82)
        Rt = np.zeros((4, 4)); Rt[:3, :3] = R.transpose(); Rt[:3, 3] = T; Rt[3, 3] = 1.0
        W2C = np.linalg.inv(Rt)
        entries.append({"id": i, "img_name": f"img{i}", "width": W, "height": H, "position": W2C[:3, 3].tolist(),
                        "rotation": [r.tolist() for r in W2C[:3, :3]], "fy": gu.fov2focal(fovy, H), "fx": gu.fov2focal(fovx, W)})
        out[f"R{i}"] = R; out[f"T{i}"] = T; out[f"fov{i}"] = np.array([fovx, fovy])
        out[f"wv{i}"] = wv.numpy(); out[f"full{i}"] = full.numpy(); out[f"center{i}"] = center.numpy()
    out["json"] = np.frombuffer(json.dumps(entries).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cameras_ref.npz"), **out)
    print("written tests/golden/cameras_ref.npz")


if __name__ == "__main__":
    main()
