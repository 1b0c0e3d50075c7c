"""CPU: the numpy restatement of the training-step elementwise work (oracle/step_oracle.py) against the dependency the reference
itself runs — torch.optim.Adam (scene/gaussian_model.py:289) and torch autograd through exp / normalize / sigmoid
(scene/gaussian_model.py:40-60) — plus the host-side behaviour of fovgs.optim.Adam that needs no GPU."""
import numpy as np
import pytest
import torch

import step_oracle as so


def _raw(P, seed):
    r = np.random.default_rng(seed)
    return (r.normal(-3, 1, (P, 3)).astype(np.float32), r.normal(0, 1, (P, 4)).astype(np.float32),
            r.normal(0, 2, (P, 1)).astype(np.float32))


def test_activations_match_torch():
    rs, rr, ro = _raw(2000, 0)
    rr[5] = 0.0                                                       # normalize's eps clamp
    s, q, o = so.activate(rs, rr, ro)
    np.testing.assert_allclose(s, torch.exp(torch.from_numpy(rs)).numpy(), rtol=2e-7)
    np.testing.assert_allclose(q, torch.nn.functional.normalize(torch.from_numpy(rr)).numpy(), rtol=3e-7, atol=1e-9)
    np.testing.assert_allclose(o, torch.sigmoid(torch.from_numpy(ro)).numpy(), rtol=3e-7)
    assert np.all(q[5] == 0)


def test_activation_backward_matches_autograd():
    rs, rr, ro = _raw(2000, 1)
    t = [torch.from_numpy(x).requires_grad_(True) for x in (rs, rr, ro)]
    s, q, o = torch.exp(t[0]), torch.nn.functional.normalize(t[1]), torch.sigmoid(t[2])
    r = np.random.default_rng(2)
    ds, dq, do = (r.normal(size=x.shape).astype(np.float32) for x in (rs, rr, ro))
    torch.autograd.backward([s, q, o], [torch.from_numpy(ds), torch.from_numpy(dq), torch.from_numpy(do)])
    g = so.activate_backward(rr, s.detach().numpy(), o.detach().numpy(), ds, dq, do)
    for a, b in zip(g, t):
        np.testing.assert_allclose(a, b.grad.numpy(), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("betas,eps", [((0.9, 0.999), 1e-15), ((0.4, 0.99), 1e-8)])
def test_adam_matches_torch(betas, eps):
    r = np.random.default_rng(3)
    shapes = [(257, 3), (257, 1, 3), (257, 15, 3), (257, 1)]
    lrs = [1.6e-4, 2.5e-3, 1.25e-4, 0.05]
    ps = [torch.nn.Parameter(torch.from_numpy(r.normal(size=s).astype(np.float32))) for s in shapes]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs)], lr=0.0, eps=eps, betas=betas, foreach=False)
    mine = [(p.detach().numpy().copy(), np.zeros(s, np.float32), np.zeros(s, np.float32)) for p, s in zip(ps, shapes)]
    for step in range(1, 8):
        for p, (w, m, v), lr in zip(ps, mine, lrs):
            g = (r.normal(size=w.shape) * 10.0 ** r.integers(-6, 1)).astype(np.float32)
            g[r.random(w.shape) < 0.3] = 0.0                            # Gaussians outside the view get exact zeros
            p.grad = torch.from_numpy(g.copy())
            so.adam_step(w, g, m, v, step, lr, betas[0], betas[1], eps)
        opt.step()
    for p, (w, m, v) in zip(ps, mine):
        st = opt.state[p]
        # fp32 tolerance: the lerp cancels (m + w*(g - m)), so a last-place difference of the FMA contraction shows as
        # 1e-7 of the array's scale, not of the element
        np.testing.assert_allclose(m, st["exp_avg"].numpy(), rtol=1e-6, atol=1e-7 * np.abs(m).max())
        np.testing.assert_allclose(v, st["exp_avg_sq"].numpy(), rtol=1e-6, atol=1e-7 * np.abs(v).max())
        np.testing.assert_allclose(w, p.detach().numpy(), rtol=1e-6, atol=1e-6)


def test_fused_adam_host_behaviour():
    """No CPU fallback, the reference's configuration only, torch.optim.Adam's state layout."""
    from fovgs import optim
    p = torch.nn.Parameter(torch.zeros(4, 3))
    o = optim.Adam([{"params": [p], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert o.param_groups[0]["name"] == "xyz" and o.param_groups[0]["eps"] == 1e-15 and o.param_groups[0]["betas"] == (0.9, 0.999)
    o.step()                                                           # no gradients: nothing to do, like torch
    assert len(o.state) == 0
    p.grad = torch.ones(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        o.step()
    with pytest.raises(NotImplementedError):
        optim.Adam([p], weight_decay=0.1)
    with pytest.raises(ValueError):
        optim.Adam([p], betas=(1.0, 0.999))
    ref = torch.optim.Adam([{"params": [p], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    o2 = optim.Adam([{"params": [p], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert set(o2.state_dict()["param_groups"][0]) >= {"lr", "betas", "eps", "name", "params"}
    o2.load_state_dict(ref.state_dict())                               # interchangeable state dicts


def test_oracle_fine_tuning_loop_descends():
    """The CPU restatements chained as eff_finetune.py:95-147 chains the real thing — activations, pcheck_obb_sum forward, L1,
    backward, activation backward, Adam — from a perturbed model towards the render of the unperturbed one: the loss must fall
    (a sign error anywhere in the gradient chain makes it rise).  The GPU twin is tests/test_gpu_step.py."""
    import oracle
    from fovgs import synth
    scene = synth.make_scene_cube(1500, 3)
    cam = synth.look_at_camera(96, 80, 60.0, (0.0, 0.0, -4.0))
    op = np.clip(scene["opacity"].astype(np.float64), 1e-4, 1 - 1e-4)
    raw0 = {"xyz": scene["means3D"].copy(), "f_dc": scene["shs"][:, :1].copy(), "f_rest": scene["shs"][:, 1:].copy(),
            "opacity": np.log(op / (1 - op)).astype(np.float32), "scaling": np.log(scene["scales"]),
            "rotation": scene["rotations"].copy()}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 5e-3, "rotation": 1e-3}

    def model(p):
        s, q, o = so.activate(p["scaling"], p["rotation"], p["opacity"])
        return {"means3D": p["xyz"], "scales": s, "rotations": q, "opacity": o, "sh_degree": 3,
                "shs": np.ascontiguousarray(np.concatenate([p["f_dc"], p["f_rest"]], 1))}

    target = oracle.forward_ps1(model(raw0), cam, "sum")["color"]
    r = np.random.default_rng(4)
    p = {k: v.copy() for k, v in raw0.items()}
    p["f_dc"] = (raw0["f_dc"] + r.normal(0, 0.3, raw0["f_dc"].shape)).astype(np.float32)
    p["opacity"] = (raw0["opacity"] + r.normal(0, 1.0, raw0["opacity"].shape)).astype(np.float32)
    m = {k: np.zeros_like(x) for k, x in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    losses = []
    for it in range(1, 16):
        sc = model(p)
        o = oracle.forward_ps1(sc, cam, "sum")
        d = o["color"] - target
        losses.append(float(np.abs(d).mean()))
        g = oracle.backward_ps1(sc, cam, o, (np.sign(d) / d.size).astype(np.float32))
        gs, gq, go = so.activate_backward(p["rotation"], sc["scales"], sc["opacity"], g["dL_dscales"], g["dL_drotations"],
                                          g["dL_dopacity"].reshape(-1, 1))
        grads = {"xyz": g["dL_dmeans3D"], "f_dc": g["dL_dsh"][:, :1], "f_rest": g["dL_dsh"][:, 1:], "opacity": go,
                 "scaling": gs, "rotation": gq}
        for k in p:
            so.adam_step(p[k], np.ascontiguousarray(grads[k], dtype=np.float32), m[k], v[k], it, lrs[k], eps=1e-15)
    assert losses[-1] < 0.6 * losses[0], losses
    assert all(b < a * 1.02 for a, b in zip(losses, losses[1:])), losses
