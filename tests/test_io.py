"""Input formats: dependency-free PLY reader/writer, cameras.json, compose_models tensors.
Pins: camera matrices against golden values computed by the reference's own utils/graphics_utils.py functions
(tools/make_io_golden.py); activations and compose rules against independent torch restatements of
scene/gaussian_model.py:40-60 and compose_models.py:39-80."""
import json
import os

import numpy as np
import pytest
import torch

from fovgs import io, synth


def _raw(P, seed, with_index=None):
    rng = np.random.default_rng(seed)
    raw = {"xyz": rng.normal(size=(P, 3)).astype(np.float32), "features_dc": rng.normal(size=(P, 1, 3)).astype(np.float32),
           "features_rest": rng.normal(size=(P, 15, 3)).astype(np.float32), "opacity": rng.normal(size=(P, 1)).astype(np.float32),
           "scaling": rng.normal(-3, 1, size=(P, 3)).astype(np.float32), "rotation": rng.normal(size=(P, 4)).astype(np.float32),
           "sh_degree": 3}
    if with_index is not None:
        raw["indexes"] = np.asarray(with_index, np.int32).reshape(P, 1)
    return raw


def test_ply_round_trip_and_header_layout(tmp_path):
    raw = _raw(257, 0)
    p = str(tmp_path / "pc" / "point_cloud.ply")
    io.write_ply(p, raw)
    head = open(p, "rb").read(400).decode("latin1")
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 257\nproperty float x\nproperty float y\n")
    v = io.read_ply(p)
    names = list(v)
    assert names[:6] == ["x", "y", "z", "nx", "ny", "nz"] and names[6:9] == ["f_dc_0", "f_dc_1", "f_dc_2"]
    assert names[9] == "f_rest_0" and names[9 + 45] == "opacity" and names[-4:] == ["rot_0", "rot_1", "rot_2", "rot_3"]
    back = io.raw_model_from_ply(p)
    for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation"):
        assert np.array_equal(back[k], raw[k]), k
    # f_rest is channel-major on disk: f_rest_0..14 = red coefficients (save_ply: transpose(1,2).flatten)
    assert np.array_equal(v["f_rest_1"], raw["features_rest"][:, 1, 0]) and np.array_equal(v["f_rest_15"], raw["features_rest"][:, 0, 1])


def test_indexed_ply_and_ascii_ply(tmp_path):
    raw = _raw(40, 1, with_index=np.arange(40)[::-1])
    p = str(tmp_path / "idx.ply")
    io.write_ply(p, raw, with_index=True)
    back = io.raw_model_from_ply(p)
    assert back["indexes"].dtype == np.int32 and np.array_equal(back["indexes"], raw["indexes"])
    # the same content as ascii
    v = io.read_ply(p)
    pa = str(tmp_path / "ascii.ply")
    with open(pa, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by a test\nelement vertex 40\n")
        for k in v:
            f.write(f"property {'int' if k == 'index' else 'float'} {k}\n")
        f.write("end_header\n")
        for i in range(40):
            f.write(" ".join(repr(float(v[k][i])) if k != "index" else str(int(v[k][i])) for k in v) + "\n")
    va = io.read_ply(pa)
    for k in v:
        assert np.array_equal(va[k], v[k]), k


def test_truncated_ply_is_rejected(tmp_path):
    raw = _raw(10, 2)
    p = str(tmp_path / "t.ply")
    io.write_ply(p, raw)
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-100])
    with pytest.raises(ValueError, match="expected 10 records"):
        io.read_ply(p)


def test_activations_match_torch():
    raw = _raw(1000, 3)
    m = io.activate(raw)
    assert np.allclose(m["scales"], torch.exp(torch.from_numpy(raw["scaling"])).numpy(), rtol=1e-6)
    assert np.allclose(m["opacity"], torch.sigmoid(torch.from_numpy(raw["opacity"])).numpy(), atol=1e-7)
    assert np.allclose(m["rotations"], torch.nn.functional.normalize(torch.from_numpy(raw["rotation"])).numpy(), atol=1e-6)
    assert m["shs"].shape == (1000, 16, 3) and np.array_equal(m["shs"][:, 0], raw["features_dc"][:, 0])


def test_compose_levels_follows_compose_models():
    P = 500
    rng = np.random.default_rng(4)
    lv = [_raw(P, 10)]
    idxs = [np.arange(P)]
    for i in range(1, 4):
        keep = np.sort(rng.choice(idxs[-1], size=len(idxs[-1]) // 2, replace=False))   # level i is a subset of level i-1
        idxs.append(keep)
        lv.append(_raw(len(keep), 10 + i, with_index=keep))
    out = io.compose_levels(lv)
    # independent restatement with torch indexing (compose_models.py:47-72)
    act = [io.activate(m) for m in lv]
    shs_dcs = torch.zeros((P, 4, 3)); highest = torch.zeros((P, 1)); opac = torch.ones((P, 4))
    shs_dcs[:, 0, :] = torch.from_numpy(act[0]["shs"][:, 0, :]); opac[:, 0] = torch.from_numpy(act[0]["opacity"][:, 0])
    for i in range(1, 4):
        ind = torch.from_numpy(lv[i]["indexes"]).long()
        shs_dcs[:, i, :] = shs_dcs[:, i - 1, :]
        shs_dcs[ind, i, :] = torch.from_numpy(act[i]["shs"][:, 0, :]).unsqueeze(1)
        opac[:, i] = opac[:, i - 1]
        opac[ind, i] = torch.from_numpy(act[i]["opacity"][:, 0]).unsqueeze(1)
        highest[ind] = i
    assert np.array_equal(out["shs_dcs"], shs_dcs.numpy()) and np.array_equal(out["opacities4"], opac.numpy())
    assert np.array_equal(out["highest_levels"], highest.numpy())
    assert out["shs_rest"].shape == (P, 15, 3) and out["highest_levels"].max() == 3


def test_camera_json_matches_reference_matrices(golden_dir):
    g = np.load(os.path.join(golden_dir, "cameras_ref.npz"))
    entries = json.loads(bytes(g["json"]).decode())
    for i, e in enumerate(entries):
        cam = io.camera_from_json_entry(e)
        assert np.allclose(cam["viewmatrix"], g[f"wv{i}"], atol=2e-6), i
        assert np.allclose(cam["projmatrix"], g[f"full{i}"], rtol=2e-6, atol=2e-5), i
        assert np.allclose(cam["campos"], g[f"center{i}"], atol=1e-5), i
        assert abs(cam["FoVx"] - g[f"fov{i}"][0]) < 1e-9 and abs(cam["FoVy"] - g[f"fov{i}"][1]) < 1e-9
        direct = io.camera_from_rt(g[f"R{i}"], g[f"T{i}"], g[f"fov{i}"][0], g[f"fov{i}"][1], e["width"], e["height"])
        assert np.array_equal(direct["viewmatrix"], g[f"wv{i}"]) and np.allclose(direct["projmatrix"], g[f"full{i}"], rtol=1e-6, atol=1e-6)
        back = io.camera_to_json_entry(i, g[f"R{i}"], g[f"T{i}"], g[f"fov{i}"][0], g[f"fov{i}"][1], e["width"], e["height"], e["img_name"])
        assert np.allclose(back["position"], e["position"]) and np.allclose(back["rotation"], e["rotation"]) and abs(back["fx"] - e["fx"]) < 1e-6


def test_synth_cameras_agree_with_io_camera_math():
    """synth.look_at_camera restates the same formulas: R/T -> identical matrices through io.camera_from_rt."""
    c = synth.ring_cameras(30, 640, 360)[7]
    wv = c["viewmatrix"].astype(np.float64)
    R = wv[:3, :3]                  # wv = Rt^T, Rt[:3,:3] = R^T  ->  wv[:3,:3] = R
    T = wv[3, :3]
    d = io.camera_from_rt(R, T, c["FoVx"], c["FoVy"], 640, 360)
    assert np.allclose(d["viewmatrix"], c["viewmatrix"], atol=1e-6) and np.allclose(d["projmatrix"], c["projmatrix"], atol=1e-5)
    assert np.allclose(d["campos"], c["campos"], atol=1e-5)


def test_composed_tensors_round_trip_as_the_reference_saves_them(tmp_path):
    """compose_models.py:75-80 / render_compose_gazes_fps.py:85-96: three torch.save files next to the model."""
    import torch
    from fovgs import synth
    f = synth.add_foveation(synth.make_scene_cube(300, 2))
    io.save_composed(str(tmp_path / "composed_4_4"), f)
    for name, shape in (("highest_levels.pt", (300, 1)), ("shs_dcs.pt", (300, 4, 3)), ("opacities.pt", (300, 4))):
        t = torch.load(str(tmp_path / "composed_4_4" / name))
        assert isinstance(t, torch.Tensor) and tuple(t.shape) == shape and t.dtype == torch.float32 and not t.is_cuda
    g = io.load_composed(str(tmp_path / "composed_4_4"))
    for k in ("highest_levels", "shs_dcs", "opacities4"):
        assert np.array_equal(g[k], f[k])
    torch.save(torch.zeros(299, 1), str(tmp_path / "composed_4_4" / "highest_levels.pt"))
    with pytest.raises(ValueError):
        io.load_composed(str(tmp_path / "composed_4_4"))


def test_smfr_levels_nest_like_gen_naive_FR():
    """gen_naive_FR.py:33-59: level i keeps the first n_i entries of level i-1's shuffled subset."""
    hl = io.smfr_levels([1000, 400, 200, 150], seed=0)
    assert hl.shape == (1000,) and hl.dtype == np.float32
    assert [(hl >= i).sum() for i in range(4)] == [1000, 400, 200, 150]
    assert not np.array_equal(hl, io.smfr_levels([1000, 400, 200, 150], seed=1))
    with pytest.raises(ValueError):
        io.smfr_levels([100, 50, 60])


def test_load_foveated_model_reads_the_layout_of_the_fps_script(tmp_path):
    """render_compose_gazes_fps.py:80-90: <base>/1_PS1_<L>_<S>/point_cloud/iteration_55000/point_cloud.ply + <base>/composed_<L>_<S>/*.pt."""
    from fovgs import synth
    sc = synth.make_scene_cube(200, 6)
    op = np.clip(sc["opacity"].astype(np.float64), 1e-4, 1 - 1e-4)
    raw = {"xyz": sc["means3D"], "features_dc": sc["shs"][:, :1], "features_rest": sc["shs"][:, 1:],
           "opacity": np.log(op / (1 - op)).astype(np.float32), "scaling": np.log(sc["scales"]), "rotation": sc["rotations"],
           "sh_degree": 3}
    d = tmp_path / "1_PS1_4_4" / "point_cloud" / "iteration_55000"
    d.mkdir(parents=True)
    io.write_ply(str(d / "point_cloud.ply"), raw)
    f = synth.add_foveation(sc)
    io.save_composed(str(tmp_path / "composed_4_4"), f)
    m = io.load_foveated_model(str(tmp_path), 4, 4)
    assert np.array_equal(m["means3D"], sc["means3D"]) and np.array_equal(m["shs_rest"], f["shs_rest"])
    assert np.allclose(m["scales"], sc["scales"], rtol=1e-6) and np.allclose(m["rotations"], sc["rotations"], atol=1e-6)
    for k in ("highest_levels", "shs_dcs", "opacities4"):
        assert np.array_equal(m[k], f[k])
    io.save_composed(str(tmp_path / "composed_4_4"), synth.add_foveation(synth.make_scene_cube(199, 6)))
    with pytest.raises(ValueError):
        io.load_foveated_model(str(tmp_path), 4, 4)
