"""CPU, only where /root/reference exists (this container): the reference's OWN render adapters — gaussian_renderer_fov,
gaussian_renderer_fov_naive, gaussian_renderer_fov_mmfr, gaussian_renderer (+ gaussian_wrapper) — imported UNMODIFIED with
`fov-3dgs_b200/` first on sys.path, called on duck-typed model / camera objects, with libfovgs.so replaced by a stub that records
the argument structs it receives.  This is SURVEY.md §8(b)'s "must run those files unchanged" checked end to end on the host
side: every keyword the adapters pass is accepted, every entry point receives the tensors of the right role and size, the
returned dict has the reference's shape, and the training variant's backward reaches `fovgs_backward_ps1`.
What is faked is the environment, never the reference's code: no GPU here, so `device="cuda"` allocations land on the CPU, and the
missing third-party `plyfile` module (imported by scene/dataset_readers.py, unused on this path) is an empty stand-in."""
import ctypes as C
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference/fov3dgs"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this machine")


class _Recorder:
    """Stands in for the ctypes library: remembers a copy of every args struct, returns success."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*a):
            rec = {"fn": name}
            if a and hasattr(a[0], "_obj"):                       # C.byref(struct)
                st = a[0]._obj
                rec["args"] = {f: getattr(st, f) for f, _ in st._fields_ if f != "cam"}
                if hasattr(st, "cam"):
                    rec["cam"] = {f: getattr(st.cam, f) for f, _ in st.cam._fields_}
            else:
                rec["raw"] = a
            self.calls.append(rec)
            return 1 << 16 if name == "fovgs_workspace_bytes" else 0
        return fn


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass


@pytest.fixture
def env(monkeypatch):
    from fovgs import ops
    rec = _Recorder()
    monkeypatch.setattr(ops, "lib", lambda: rec)
    monkeypatch.setattr(ops, "_pool", ops._Pool())
    monkeypatch.setattr(ops, "_train_capacity_hint", {})
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())

    def prep(t, name, device, dtype=torch.float32, optional=False):   # ops._prep minus the is_cuda requirement
        if t is None or (isinstance(t, torch.Tensor) and t.numel() == 0):
            if optional:
                return None
            raise RuntimeError(f"{name} must be a non-empty tensor")
        return (t if t.dtype == dtype else t.to(dtype)).contiguous()

    monkeypatch.setattr(ops, "_prep", prep)
    real_zeros_like = torch.zeros_like

    def zeros_like(t, *a, **k):
        k.pop("device", None)                                          # the adapters ask for device="cuda"
        return real_zeros_like(t, *a, **k)

    monkeypatch.setattr(torch, "zeros_like", zeros_like)
    if "plyfile" not in sys.modules:
        ply = types.ModuleType("plyfile")
        ply.PlyData = ply.PlyElement = object
        monkeypatch.setitem(sys.modules, "plyfile", ply)
    monkeypatch.syspath_prepend(REF)
    for m in ("gaussian_renderer_fov", "gaussian_renderer_fov_naive", "gaussian_renderer_fov_mmfr", "gaussian_renderer",
              "gaussian_wrapper"):
        sys.modules.pop(m, None)
    yield rec
    for m in [k for k in sys.modules if k.split(".")[0] in ("gaussian_renderer_fov", "gaussian_renderer_fov_naive",
                                                             "gaussian_renderer_fov_mmfr", "gaussian_renderer", "gaussian_wrapper",
                                                             "scene", "utils", "arguments")]:
        sys.modules.pop(m, None)


P, W, H = 50, 96, 64


def _model(requires_grad=False):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g).requires_grad_(requires_grad)
    feats = r(P, 16, 3)
    return types.SimpleNamespace(get_xyz=r(P, 3), get_opacity=r(P, 1), get_scaling=r(P, 3), get_rotation=r(P, 4),
                                 get_features=feats, get_rest_features=feats[:, 1:], get_features_detach_rest=feats,
                                 active_sh_degree=3)


def _camera():
    return types.SimpleNamespace(FoVx=1.0, FoVy=0.7, image_height=H, image_width=W, world_view_transform=torch.eye(4),
                                 full_proj_transform=torch.eye(4), camera_center=torch.zeros(3))


def _last(rec, fn):
    calls = [c for c in rec.calls if c["fn"] == fn]
    assert calls, f"{fn} was never reached; calls: {[c['fn'] for c in rec.calls]}"
    return calls[-1]


def test_fov_adapter_runs_unchanged(env):
    mod = importlib.import_module("gaussian_renderer_fov")
    assert mod.__file__.startswith(REF)
    pc = _model()
    gaze = torch.tensor([0.25, 0.75])
    with torch.no_grad():
        out = mod.render(_camera(), pc, torch.zeros(3), alpha=0.05, gazeArray=gaze, blending=True,
                         highest_levels=torch.zeros(P, 1), shs_dcs=torch.rand(P, 4, 3), opacities=torch.rand(P, 4))
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii"}
    assert tuple(out["render"].shape) == (3, H, W) and tuple(out["radii"].shape) == (P,) and out["visibility_filter"].dtype == torch.bool
    c = _last(env, "fovgs_forward_fov")
    a = c["args"]
    assert a["P"] == P and a["M_rest"] == 15 and abs(a["alpha"] - 0.05) < 1e-7 and a["blending"] == 1
    assert c["cam"]["image_width"] == W and c["cam"]["image_height"] == H and c["cam"]["sh_degree"] == 3
    assert abs(c["cam"]["tanfovx"] - np.tan(0.5)) < 1e-6 and abs(c["cam"]["tanfovy"] - np.tan(0.35)) < 1e-6
    for k in ("means3D", "opacities", "scales", "rotations", "shs_rest", "shs_dcs", "highest_levels", "gaze", "out_color", "radii",
              "workspace"):
        assert a[k], k
    assert a["means3D"] == pc.get_xyz.data_ptr() and a["scales"] == pc.get_scaling.data_ptr()        # borrowed, not copied
    assert a["out_color"] == out["render"].data_ptr() and a["gaze"] == gaze.data_ptr()


def test_smfr_adapter_runs_unchanged(env):
    mod = importlib.import_module("gaussian_renderer_fov_naive")
    pc = _model()
    with torch.no_grad():
        out = mod.render(_camera(), pc, torch.zeros(3), alpha=0.05, gazeArray=torch.tensor([0.5, 0.5]), blending=True,
                         highest_levels=torch.zeros(P))
    assert tuple(out["render"].shape) == (3, H, W)
    a = _last(env, "fovgs_forward_smfr")["args"]
    assert a["P"] == P and a["M"] == 16 and a["shs"] == pc.get_features.data_ptr() and a["highest_levels"]


def test_mmfr_adapter_calls_one_level_at_a_time(env):
    mod = importlib.import_module("gaussian_renderer_fov_mmfr")
    models = [_model() for _ in range(4)]
    with torch.no_grad():
        out = mod.render(_camera(), torch.zeros(3), alpha=0.05, gazeArray=torch.tensor([0.5, 0.5]), blending=True,
                         multi_gs=models, layer_num=4)
    assert tuple(out["render"].shape) == (3, H, W)
    calls = [c["args"] for c in env.calls if c["fn"] == "fovgs_forward_mmfr"]
    assert [c["cur_level"] for c in calls] == [0, 1, 2, 3]
    assert [c["means3D"] for c in calls] == [m.get_xyz.data_ptr() for m in models]


@pytest.mark.parametrize("cuda_type,mode,n_out", [("pcheck_obb", 0, 4), ("pcheck_obb_sum", 1, 6), ("pcheck_obb_max", 2, 6),
                                                  ("original", 4, 4)])   # gaussian_wrapper.py:11: the stock rasterizer
def test_ps1_adapter_and_wrapper_run_unchanged(env, cuda_type, mode, n_out):
    mod = importlib.import_module("gaussian_renderer")
    pc = _model()
    pipe = types.SimpleNamespace(debug=False)
    with torch.no_grad():
        out = mod.render(_camera(), pc, pipe, torch.zeros(3), cuda_type=cuda_type)
    assert len(out) == n_out and tuple(out["render"].shape) == (3, H, W)
    if n_out == 6:
        assert tuple(out["gs_count"].shape) == (P,) and tuple(out["contribs"].shape) == (P,)
    c = _last(env, "fovgs_forward_ps1")
    assert c["args"]["P"] == P and c["args"]["M"] == 16 and c["args"]["mode"] == mode and c["args"]["shs"] == pc.get_features.data_ptr()


def test_loss_weighted_adapter_passes_the_loss_map(env):
    mod = importlib.import_module("gaussian_renderer")
    lm = torch.rand(H, W)
    with torch.no_grad():
        out = mod.render(_camera(), _model(), types.SimpleNamespace(debug=False), torch.zeros(3),
                         cuda_type="pcheck_obb_loss_weighted_max_count", loss_map=lm)
    assert "contribs" in out
    a = _last(env, "fovgs_forward_ps1")["args"]
    assert a["mode"] == 3 and a["loss_map"] == lm.data_ptr()


def test_training_step_through_the_adapter_reaches_the_backward_entry(env):
    """eff_finetune.py:107-127: render(..., cuda_type="pcheck_obb_sum"), loss, loss.backward()."""
    mod = importlib.import_module("gaussian_renderer")
    pc = _model(requires_grad=True)
    out = mod.render(_camera(), pc, types.SimpleNamespace(debug=False), torch.zeros(3), cuda_type="pcheck_obb_sum")
    assert out["render"].requires_grad
    out["render"].abs().mean().backward()
    b = _last(env, "fovgs_backward_ps1")["args"]
    assert b["P"] == P and b["dL_dout_color"] and b["dL_dmeans3D"] and b["dL_dsh"] and b["workspace"]
    for t in (pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation, pc.get_features):
        assert t.grad is not None and t.grad.shape == t.shape
    assert out["viewspace_points"].grad is not None                     # densification statistics read this (train.py convention)
