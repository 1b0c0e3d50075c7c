#!/usr/bin/env python
"""One whole fine-tuning iteration (BASELINE config 5; eff_finetune.py:95-147) on the bench scene, both arms on the same GPU:

  activations -> cat(features_dc, features_rest) -> pcheck_obb_sum rasterizer -> 0.8*L1 + 0.2*(1 - SSIM) -> backward -> Adam

  ours      : fovgs.ops.activate, diff_gaussian_rasterization_pcheck_obb_sum (this repo), fovgs.optim.Adam
  reference : torch.exp / F.normalize / torch.sigmoid, the unmodified reference CUDA (oracle/_ref/ref_sum_C) under an autograd
              Function, torch.optim.Adam(l, lr=0.0, eps=1e-15)            (only when oracle/_ref is built)
The loss (utils/loss_utils.py: l1_loss + 11x11 Gaussian-window SSIM) is the same torch code in both arms.  CUDA events around
whole iterations, 3 warm-ups, `--iters` timed; one JSON line.

  python tools/train_step.py [--size big] [--iters 10]
"""
import argparse, json, os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
from fovgs import ops, optim  # noqa: E402

LRS = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 5e-3, "rotation": 1e-3}


def window(device):
    g = torch.tensor([np.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    return (g @ g.t()).expand(3, 1, 11, 11).contiguous().to(device)


def ssim(a, b, w):
    mu1, mu2 = F.conv2d(a, w, padding=5, groups=3), F.conv2d(b, w, padding=5, groups=3)
    s1 = F.conv2d(a * a, w, padding=5, groups=3) - mu1 * mu1
    s2 = F.conv2d(b * b, w, padding=5, groups=3) - mu2 * mu2
    s12 = F.conv2d(a * b, w, padding=5, groups=3) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()


def raw_params(scene, dev):
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    op = np.clip(scene["opacity"].astype(np.float64), 1e-6, 1 - 1e-6)
    shs = scene["shs"]
    raw = {"xyz": t(scene["means3D"]), "f_dc": t(shs[:, :1]), "f_rest": t(shs[:, 1:]),
           "opacity": t(np.log(op / (1 - op)).astype(np.float32).reshape(-1, 1)),
           "scaling": t(np.log(scene["scales"])), "rotation": t(scene["rotations"])}
    return {k: torch.nn.Parameter(v.requires_grad_(True)) for k, v in raw.items()}


class RefSum(torch.autograd.Function):
    """The reference's _RasterizeGaussians (SUM/diff_gaussian_rasterization_pcheck_obb_sum/__init__.py:44-126) over its own _C."""

    @staticmethod
    def forward(ctx, mod, cam, means3D, shs, opacity, scales, rotations):
        import ref_api
        sc = {"means3D": means3D, "opacity": opacity, "scales": scales, "rotations": rotations, "shs": shs, "sh_degree": 3}
        res = ref_api.ps1_forward(mod, sc, cam)
        n, color, radii, geom, binning, img = res[:6]
        ctx.mod, ctx.cam, ctx.n = mod, cam, n
        ctx.save_for_backward(means3D, shs, scales, rotations, radii, geom, binning, img)
        return color

    @staticmethod
    def backward(ctx, grad_out):
        import ref_api
        means3D, shs, scales, rotations, radii, geom, binning, img = ctx.saved_tensors
        sc = {"means3D": means3D, "scales": scales, "rotations": rotations, "shs": shs, "sh_degree": 3}
        g = ref_api.ps1_backward(ctx.mod, sc, ctx.cam, radii, grad_out.contiguous(), geom, ctx.n, binning, img)
        # dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations
        return None, None, g[3], g[5], g[2], g[6], g[7]


def run(arm, wl, dev, iters, mod=None):
    p = raw_params(wl.scene, dev)
    groups = [{"params": [p[k]], "lr": LRS[k], "name": k} for k in LRS]
    opt = (optim.Adam if arm == "ours" else torch.optim.Adam)(groups, lr=0.0, eps=1e-15)
    cams = [bench.to_dev(c, dev) for c in wl.cams[:8]]
    bg = torch.zeros(3, device=dev)
    w = window(dev)
    gt = torch.rand(3, wl.H, wl.W, device=dev)
    import diff_gaussian_rasterization_pcheck_obb_sum as pkg
    losses = []

    def iteration(i):
        c = cams[i % len(cams)]
        if arm == "ours":
            scales, rots, opac = ops.activate(p["scaling"], p["rotation"], p["opacity"])
        else:
            scales, rots, opac = torch.exp(p["scaling"]), F.normalize(p["rotation"]), torch.sigmoid(p["opacity"])
        shs = torch.cat((p["f_dc"], p["f_rest"]), dim=1)
        if arm == "ours":
            rs = pkg.GaussianRasterizationSettings(c["image_height"], c["image_width"], c["tanfovx"], c["tanfovy"], bg, 1.0,
                                                   c["viewmatrix"], c["projmatrix"], 3, c["campos"], False, False)
            means2D = torch.zeros_like(p["xyz"], requires_grad=True)
            img = pkg.GaussianRasterizer(raster_settings=rs)(means3D=p["xyz"], means2D=means2D, shs=shs, colors_precomp=None,
                                                             opacities=opac, scales=scales, rotations=rots, cov3D_precomp=None)[0]
        else:
            img = RefSum.apply(mod, c, p["xyz"], shs, opac, scales, rots)
        loss = 0.8 * (img - gt).abs().mean() + 0.2 * (1.0 - ssim(img.unsqueeze(0), gt.unsqueeze(0), w))
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(3):
        iteration(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        losses.append(iteration(3 + i).detach())
    b.record()
    torch.cuda.synchronize()
    return {"ms_per_iteration": a.elapsed_time(b) / iters, "loss_first": float(losses[0]), "loss_last": float(losses[-1])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="big"); ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = bench.Workload(a.size)
    out = {"workload": f"train_{wl.P // 1000}k_{wl.W}x{wl.H}_pcheck_obb_sum", "iters": a.iters}
    out["ours"] = run("ours", wl, dev, a.iters)
    torch.cuda.empty_cache()
    try:
        import ref_api
        mod = ref_api.ref_module("ref_sum_C")
    except Exception:
        mod = None
    if mod is not None:
        out["reference"] = run("reference", wl, dev, a.iters, mod)
    else:
        out["reference"] = {"unavailable": "oracle/_ref/ref_sum_C not built (FOVGS_BUILD_ALL_REFS=1 python oracle/build_ref.py)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
