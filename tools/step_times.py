#!/usr/bin/env python
"""Times the training-step elementwise kernels at the bench model size against the code the reference
runs on the device: torch.exp / F.normalize / torch.sigmoid (+ autograd) and torch.optim.Adam(l, lr=0.0, eps=1e-15)
(foreach, the default on CUDA, and fused=True).  CUDA events, 3 warm-ups, 10 timed repetitions, one JSON line.

  python tools/step_times.py [--P 6000000]
"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
from fovgs import ops, optim  # noqa: E402

SHAPES = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def params(P):
    return [{"params": [torch.nn.Parameter(torch.randn((P,) + s, device="cuda"))], "lr": 1e-3, "name": n} for n, s in SHAPES.items()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=6_000_000)
    a = ap.parse_args()
    P = a.P
    out = {"P": P}
    rs, rr, ro = (torch.randn(P, k, device="cuda", requires_grad=True) for k in (3, 4, 1))
    d = [torch.randn(P, k, device="cuda") for k in (3, 4, 1)]

    def act_torch():
        s, q, o = torch.exp(rs), torch.nn.functional.normalize(rr), torch.sigmoid(ro)
        torch.autograd.backward([s, q, o], d)
        rs.grad = rr.grad = ro.grad = None

    def act_ours():
        s, q, o = ops.activate(rs, rr, ro)
        torch.autograd.backward([s, q, o], d)
        rs.grad = rr.grad = ro.grad = None

    out["activate_fwd_bwd_ms"] = {"torch": timeit(act_torch), "ours": timeit(act_ours)}
    out["activate_bytes"] = P * 8 * 4 * (2 + 4)   # forward: read + write 8 floats; backward: ~3 reads + 1 write of 8 floats
    n_el = P * 59
    out["adam_elements"] = n_el
    out["adam_alg_bytes"] = n_el * 28
    res = {}
    for name, make in (("torch_foreach", lambda l: torch.optim.Adam(l, lr=0.0, eps=1e-15)),
                       ("torch_fused", lambda l: torch.optim.Adam(l, lr=0.0, eps=1e-15, fused=True)),
                       ("ours", lambda l: optim.Adam(l, lr=0.0, eps=1e-15))):
        l = params(P)
        for g in l:
            g["params"][0].grad = torch.randn_like(g["params"][0])
        opt = make(l)
        res[name] = timeit(opt.step)
        del opt, l
        torch.cuda.empty_cache()
    out["adam_step_ms"] = res
    out["adam_ours_GBps"] = out["adam_alg_bytes"] / (res["ours"] * 1e-3) / 1e9
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        try:
            out["peaks"] = json.load(open(peaks))
        except Exception:
            pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
