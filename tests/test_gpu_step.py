"""GPU (pytest -m gpu): the training-step elementwise kernels (fovgs_activate_forward/_backward, fovgs_adam_step) through the
C-ABI against the numpy oracle (oracle/step_oracle.py, pinned against torch on the CPU) and against the code the reference runs
on the device: torch.exp / F.normalize / torch.sigmoid under autograd (scene/gaussian_model.py:40-60) and
torch.optim.Adam(l, lr=0.0, eps=1e-15) (scene/gaussian_model.py:289).  Floating-point bars are written at each assert."""
import copy

import numpy as np
import pytest
import torch

import step_oracle as so
from fovgs import ops, optim

pytestmark = pytest.mark.gpu


def _raw(P, seed):
    r = np.random.default_rng(seed)
    return (r.normal(-3, 1, (P, 3)).astype(np.float32), r.normal(0, 1, (P, 4)).astype(np.float32),
            r.normal(0, 2, (P, 1)).astype(np.float32))


@pytest.mark.parametrize("P", [1, 1000, 300001])
def test_activate_forward_and_backward(P):
    rs, rr, ro = _raw(P, P)
    if P > 5:
        rr[5] = 0.0
    t = [torch.from_numpy(x).cuda().requires_grad_(True) for x in (rs, rr, ro)]
    s, q, o = ops.activate(*t)
    es, eq, eo = so.activate(rs, rr, ro)
    # forward: exp and sigmoid within 2 ulp of the numpy restatement (different libm), normalize likewise
    np.testing.assert_allclose(s.detach().cpu().numpy(), es, rtol=3e-7)
    np.testing.assert_allclose(q.detach().cpu().numpy(), eq, rtol=4e-7, atol=1e-9)
    np.testing.assert_allclose(o.detach().cpu().numpy(), eo, rtol=4e-7)
    # and against the device code the reference runs
    t2 = [x.detach().clone().requires_grad_(True) for x in t]
    s2, q2, o2 = torch.exp(t2[0]), torch.nn.functional.normalize(t2[1]), torch.sigmoid(t2[2])
    assert torch.allclose(s, s2, rtol=3e-7, atol=0) and torch.allclose(q, q2, rtol=4e-7, atol=1e-9) and torch.allclose(o, o2, rtol=4e-7, atol=0)
    r = np.random.default_rng(7)
    d = [torch.from_numpy(r.normal(size=x.shape).astype(np.float32)).cuda() for x in (rs, rr, ro)]
    torch.autograd.backward([s, q, o], d)
    torch.autograd.backward([s2, q2, o2], d)
    g = so.activate_backward(rr, es, eo, *(x.cpu().numpy() for x in d))
    for mine, ref, orc in zip(t, t2, g):
        np.testing.assert_allclose(mine.grad.cpu().numpy(), orc, rtol=2e-5, atol=2e-6)
        assert torch.allclose(mine.grad, ref.grad, rtol=2e-5, atol=2e-6)


def test_activate_partial_grads_and_errors():
    rs, rr, ro = _raw(100, 3)
    a, b, c = (torch.from_numpy(x).cuda() for x in (rs, rr, ro))
    b.requires_grad_(True)
    s, q, o = ops.activate(a, b, c)
    q.sum().backward()
    assert b.grad is not None and a.grad is None
    with pytest.raises(RuntimeError, match="dimensions"):
        ops.activate(a[:, :2], b, c)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.activate(a.cpu(), b.cpu(), c.cpu())


def _model_params(P, seed, device="cuda"):
    r = np.random.default_rng(seed)
    shapes = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, 15, 3), "opacity": (P, 1), "scaling": (P, 3), "rotation": (P, 4)}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 5e-3, "rotation": 1e-3}
    return [{"params": [torch.nn.Parameter(torch.from_numpy(r.normal(size=s).astype(np.float32)).to(device))], "lr": lrs[n], "name": n}
            for n, s in shapes.items()]


@pytest.mark.parametrize("P", [3, 1001, 40000])
def test_adam_matches_torch_and_oracle(P):
    """Six groups as scene/gaussian_model.py:279-287 builds them; P = 1001 leaves ragged tails (the scalar path), P = 3 is all tail."""
    mine, ref = _model_params(P, 11), _model_params(P, 11)
    o_mine = optim.Adam(mine, lr=0.0, eps=1e-15)
    o_ref = torch.optim.Adam(ref, lr=0.0, eps=1e-15)
    orc = [(g["params"][0].detach().cpu().numpy().copy(), np.zeros(tuple(g["params"][0].shape), np.float32),
            np.zeros(tuple(g["params"][0].shape), np.float32)) for g in mine]
    r = np.random.default_rng(12)
    for step in range(1, 7):
        for gm, gr, (w, m, v) in zip(mine, ref, orc):
            shp = tuple(gm["params"][0].shape)
            g = (r.normal(size=shp) * 10.0 ** r.integers(-6, 1)).astype(np.float32)
            g[r.random(shp) < 0.4] = 0.0
            gm["params"][0].grad = torch.from_numpy(g).cuda()
            gr["params"][0].grad = torch.from_numpy(g).cuda()
            so.adam_step(w, g, m, v, step, gm["lr"], eps=1e-15)
        o_mine.step()
        o_ref.step()
    for gm, gr, (w, m, v) in zip(mine, ref, orc):
        pm, pr = gm["params"][0], gr["params"][0]
        sm, sr = o_mine.state[pm], o_ref.state[pr]
        assert float(sm["step"]) == float(sr["step"]) == 6
        for a, b, c in ((sm["exp_avg"], sr["exp_avg"], m), (sm["exp_avg_sq"], sr["exp_avg_sq"], v)):
            scale = float(b.abs().max())
            # fp32: 1e-6 relative, plus 1e-7 of the array's scale where the moment update cancels
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7 * scale), gm["name"]
            np.testing.assert_allclose(a.cpu().numpy(), c, rtol=1e-6, atol=1e-7 * scale)
        assert torch.allclose(pm, pr, rtol=1e-6, atol=1e-6), gm["name"]
        np.testing.assert_allclose(pm.detach().cpu().numpy(), w, rtol=1e-6, atol=1e-6)


def test_adam_unaligned_views_and_missing_grads():
    """A parameter whose storage is not 16-byte aligned takes the scalar path; parameters without a gradient are skipped."""
    base = torch.randn(4 * 5000 + 1, device="cuda")
    p_un = torch.nn.Parameter(base[1:].view(5000, 4))                       # data_ptr offset by 4 bytes
    p_no = torch.nn.Parameter(torch.randn(100, 3, device="cuda"))
    q_un, q_no = torch.nn.Parameter(p_un.detach().clone()), torch.nn.Parameter(p_no.detach().clone())
    a = optim.Adam([{"params": [p_un], "lr": 1e-2}, {"params": [p_no], "lr": 1e-2}], lr=0.0, eps=1e-15)
    b = torch.optim.Adam([{"params": [q_un], "lr": 1e-2}, {"params": [q_no], "lr": 1e-2}], lr=0.0, eps=1e-15)
    for _ in range(3):
        g = torch.randn(5000, 4, device="cuda")
        p_un.grad, q_un.grad = g.clone(), g.clone()
        a.step(); b.step()
    assert torch.allclose(p_un, q_un, rtol=1e-6, atol=1e-6)
    assert torch.equal(p_no, q_no) and p_no not in a.state


def test_adam_survives_the_reference_optimizer_surgery():
    """_prune_optimizer (scene/gaussian_model.py:624-640) masks the parameter and both moments and re-keys the state."""
    mine, ref = _model_params(500, 5), _model_params(500, 5)
    o_mine, o_ref = optim.Adam(mine, lr=0.0, eps=1e-15), torch.optim.Adam(ref, lr=0.0, eps=1e-15)

    def step_all(seed):
        r = torch.Generator(device="cuda").manual_seed(seed)
        for gm, gr in zip(o_mine.param_groups, o_ref.param_groups):
            g = torch.randn(gm["params"][0].shape, device="cuda", generator=r)
            gm["params"][0].grad, gr["params"][0].grad = g.clone(), g.clone()
        o_mine.step(); o_ref.step()

    def prune(opt, mask):
        for group in opt.param_groups:
            st = opt.state.get(group["params"][0], None)
            st["exp_avg"] = st["exp_avg"][mask]
            st["exp_avg_sq"] = st["exp_avg_sq"][mask]
            del opt.state[group["params"][0]]
            group["params"][0] = torch.nn.Parameter(group["params"][0][mask].requires_grad_(True))
            opt.state[group["params"][0]] = st

    step_all(0); step_all(1)
    mask = torch.rand(500, device="cuda") > 0.3
    prune(o_mine, mask); prune(o_ref, mask)
    step_all(2)
    for gm, gr in zip(o_mine.param_groups, o_ref.param_groups):
        assert gm["params"][0].shape == gr["params"][0].shape
        assert torch.allclose(gm["params"][0], gr["params"][0], rtol=1e-6, atol=1e-6)
    # state dicts are interchangeable with torch.optim.Adam's (deep copies: Optimizer.load_state_dict keeps the very tensors
    # it is handed, and two optimizers sharing their moments would each update them)
    o_ref.load_state_dict(copy.deepcopy(o_mine.state_dict()))
    o_mine.load_state_dict(copy.deepcopy(o_ref.state_dict()))
    step_all(3)
    for gm, gr in zip(o_mine.param_groups, o_ref.param_groups):
        assert torch.allclose(gm["params"][0], gr["params"][0], rtol=1e-6, atol=1e-6)


def test_fine_tuning_loop_reduces_the_loss_like_the_torch_side():
    """eff_finetune.py:95-147 in miniature: activations -> pcheck_obb_sum rasterizer -> L1 -> backward -> Adam, 40 iterations
    from a perturbed model towards the render of the unperturbed one.  Arm A uses fovgs.ops.activate + fovgs.optim.Adam,
    arm B torch.exp / F.normalize / torch.sigmoid + torch.optim.Adam around the same rasterizer.  Both must reduce the loss,
    and end within 5 % of each other (parameter-wise equality is not a property of Adam with eps = 1e-15: an update is
    lr * sign(g) for the first step, and the rasterizer's atomics make g's last bits run-dependent)."""
    import diff_gaussian_rasterization_pcheck_obb_sum as pkg
    from fovgs import synth
    scene = synth.make_scene_cube(3000, 3)
    cam = synth.look_at_camera(160, 128, 60.0, (0.0, 0.0, -4.0))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    bg = torch.zeros(3, device="cuda")
    rs = pkg.GaussianRasterizationSettings(cam["image_height"], cam["image_width"], cam["tanfovx"], cam["tanfovy"], bg, 1.0,
                                           t(cam["viewmatrix"]), t(cam["projmatrix"]), 3, t(cam["campos"]), False, False)
    rast = pkg.GaussianRasterizer(raster_settings=rs)
    op = np.clip(scene["opacity"].astype(np.float64), 1e-4, 1 - 1e-4)
    raw0 = {"xyz": scene["means3D"], "f_dc": scene["shs"][:, :1], "f_rest": scene["shs"][:, 1:],
            "opacity": np.log(op / (1 - op)).astype(np.float32), "scaling": np.log(scene["scales"]), "rotation": scene["rotations"]}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 5e-3, "rotation": 1e-3}

    def render(p, ours):
        if ours:
            s, q, o = ops.activate(p["scaling"], p["rotation"], p["opacity"])
        else:
            s, q, o = torch.exp(p["scaling"]), torch.nn.functional.normalize(p["rotation"]), torch.sigmoid(p["opacity"])
        shs = torch.cat((p["f_dc"], p["f_rest"]), dim=1)
        return rast(means3D=p["xyz"], means2D=torch.zeros_like(p["xyz"], requires_grad=True), shs=shs, colors_precomp=None,
                    opacities=o, scales=s, rotations=q, cov3D_precomp=None)[0]

    with torch.no_grad():
        target = render({k: t(v) for k, v in raw0.items()}, True)
    r = np.random.default_rng(4)
    start = dict(raw0)
    start["f_dc"] = (raw0["f_dc"] + r.normal(0, 0.3, raw0["f_dc"].shape)).astype(np.float32)
    start["opacity"] = (raw0["opacity"] + r.normal(0, 1.0, raw0["opacity"].shape)).astype(np.float32)

    def run(ours):
        p = {k: torch.nn.Parameter(t(v)) for k, v in start.items()}
        groups = [{"params": [p[k]], "lr": lrs[k], "name": k} for k in lrs]
        opt = (optim.Adam if ours else torch.optim.Adam)(groups, lr=0.0, eps=1e-15)
        losses = []
        for _ in range(40):
            loss = (render(p, ours) - target).abs().mean()
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            losses.append(float(loss.detach()))
        return losses

    a, b = run(True), run(False)
    assert abs(a[0] - b[0]) <= 1e-5 * a[0]                             # same start; activations agree to an ulp
    assert a[-1] < 0.7 * a[0] and b[-1] < 0.7 * b[0], (a[0], a[-1], b[0], b[-1])
    assert abs(a[-1] - b[-1]) <= 0.05 * b[-1], (a[-1], b[-1])
