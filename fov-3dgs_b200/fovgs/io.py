"""Input formats either side of the hot path, without the reference's `plyfile` dependency.

  read_ply / write_ply        point_cloud.ply as GaussianModel.save_ply / save_ply_index / load_ply / load_ply_index
                              write and read it (scene/gaussian_model.py:303-398, 454-546): one `vertex` element with
                              x,y,z,nx,ny,nz, f_dc_0..2, f_rest_0..44, opacity, scale_0..2, rot_0..3 [, index] —
                              float32 little-endian (index int32); ascii PLY is read as well.
  model_from_ply              the raw (pre-activation) parameters -> the rasterizer's input tensors: the reference's
                              activations exp / sigmoid / normalize (scene/gaussian_model.py:40-60, 200-237) and its
                              feature layout (f_rest is stored channel-major [P,3,15] and transposed to [P,15,3]).
  compose_levels              compose_models.py:39-80: highest_levels [P,1], shs_dcs [P,L,3], opacities [P,L] of the
                              multi-level ("ours") model from the level-0 PLY and the indexed PLYs of levels 1..L-1.
  save_composed / load_composed   the three tensors compose_models.py:75-80 saves next to the model and
                              render_compose_gazes_fps.py:85-96 loads: highest_levels.pt [P,1], shs_dcs.pt [P,L,3], opacities.pt
                              [P,L] (torch.save files, float32)
  smfr_levels                 gen_naive_FR.py:33-59: the SMFR baseline's highest_levels.pth — nested random subsets of one model
  camera_from_json_entry      one entry of cameras.json (utils/camera_utils.py:62-82) -> the camera dict the rasterizer
  cameras_from_json           settings are built from (scene/cameras.py:17-57, utils/graphics_utils.py:38-71).

Everything is numpy on the host; tensors go to the device in the caller.  Nothing here is on the timed path.
"""
import json
import math
import os

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def _parse_header(f):
    magic = f.readline().strip()
    if magic != b"ply":
        raise ValueError("not a PLY file")
    fmt = None
    elements = []          # [name, count, [(prop, dtype)]]
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header is not terminated by end_header")
        tok = line.decode("ascii", "replace").strip().split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            if tok[1] == "list":
                raise ValueError("PLY list properties are not used by Gaussian point clouds")
            if tok[1] not in _PLY_TYPES:
                raise ValueError(f"unknown PLY property type {tok[1]}")
            elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
        elif tok[0] == "end_header":
            break
    if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
        raise ValueError(f"unsupported PLY format {fmt}")
    return fmt, elements


def read_ply(path):
    """Returns {property name: 1-D array} of the `vertex` element (all properties, file order preserved in dict order)."""
    with open(path, "rb") as f:
        fmt, elements = _parse_header(f)
        for name, count, props in elements:
            if fmt == "ascii":
                rows = np.loadtxt(f, max_rows=count, ndmin=2, dtype=np.float64) if count else np.zeros((0, len(props)))
                data = {p: rows[:, i].astype(t) for i, (p, t) in enumerate(props)}
            else:
                end = "<" if fmt == "binary_little_endian" else ">"
                dt = np.dtype([(p, end + t) for p, t in props])
                rec = np.fromfile(f, dtype=dt, count=count)
                if rec.shape[0] != count:
                    raise ValueError(f"PLY element {name}: expected {count} records, file holds {rec.shape[0]}")
                data = {p: np.ascontiguousarray(rec[p]).astype(t) for p, t in props}
            if name == "vertex":
                return data
    raise ValueError("PLY file has no vertex element")


def _sorted_props(d, prefix):
    names = [k for k in d if k.startswith(prefix)]
    return sorted(names, key=lambda x: int(x.split("_")[-1]))


def raw_model_from_ply(path, max_sh_degree=3):
    """The reference's load_ply / load_ply_index: RAW parameters in GaussianModel's layout:
    xyz [P,3], features_dc [P,1,3], features_rest [P,(D+1)^2-1,3], opacity [P,1] (logit), scaling [P,3] (log),
    rotation [P,4] (unnormalised), indexes [P,1] int32 when the file has an `index` property."""
    v = read_ply(path)
    P = v["x"].shape[0]
    xyz = np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float32)
    dc = np.stack([v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]], axis=1).astype(np.float32).reshape(P, 3, 1)
    rest_names = _sorted_props(v, "f_rest_")
    n_rest = (max_sh_degree + 1) ** 2 - 1
    if len(rest_names) != 3 * n_rest:
        raise ValueError(f"expected {3 * n_rest} f_rest properties for SH degree {max_sh_degree}, found {len(rest_names)}")
    rest = np.stack([v[n] for n in rest_names], axis=1).astype(np.float32).reshape(P, 3, n_rest) if n_rest else np.zeros((P, 3, 0), np.float32)
    out = {
        "xyz": xyz,
        "features_dc": np.ascontiguousarray(dc.transpose(0, 2, 1)),        # [P,1,3]
        "features_rest": np.ascontiguousarray(rest.transpose(0, 2, 1)),    # [P,n_rest,3]
        "opacity": v["opacity"].astype(np.float32).reshape(P, 1),
        "scaling": np.stack([v[n] for n in _sorted_props(v, "scale_")], axis=1).astype(np.float32),
        "rotation": np.stack([v[n] for n in _sorted_props(v, "rot")], axis=1).astype(np.float32),
        "sh_degree": int(max_sh_degree),
    }
    if "index" in v:
        out["indexes"] = v["index"].astype(np.int32).reshape(P, 1)
    return out


def write_ply(path, raw, with_index=False):
    """The reference's save_ply / save_ply_index: binary little-endian, properties in construct_list_of_attributes order."""
    P = raw["xyz"].shape[0]
    f_dc = raw["features_dc"].transpose(0, 2, 1).reshape(P, -1)
    f_rest = raw["features_rest"].transpose(0, 2, 1).reshape(P, -1)
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(f_dc.shape[1])] + [f"f_rest_{i}" for i in range(f_rest.shape[1])] + ["opacity"]
    names += [f"scale_{i}" for i in range(raw["scaling"].shape[1])] + [f"rot_{i}" for i in range(raw["rotation"].shape[1])]
    cols = np.concatenate([raw["xyz"], np.zeros_like(raw["xyz"]), f_dc, f_rest, raw["opacity"].reshape(P, 1), raw["scaling"],
                           raw["rotation"]], axis=1).astype("<f4")
    dt = [(n, "<f4") for n in names]
    if with_index:
        dt.append(("index", "<i4"))
    rec = np.empty(P, dtype=np.dtype(dt))
    for i, n in enumerate(names):
        rec[n] = cols[:, i]
    if with_index:
        rec["index"] = raw["indexes"].reshape(P).astype("<i4")
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {P}"]
        hdr += [f"property float {n}" for n in names]
        if with_index:
            hdr.append("property int index")
        hdr.append("end_header")
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        rec.tofile(f)


def activate(raw):
    """GaussianModel's activations (scene/gaussian_model.py:40-60): the tensors the rasterizer takes.
    Returns means3D, scales = exp, rotations = normalised, opacity = sigmoid, shs [P,(D+1)^2,3] (dc first), sh_degree."""
    q = raw["rotation"].astype(np.float32)
    n = np.sqrt((q.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    n = np.maximum(n, 1e-12)                                  # torch.nn.functional.normalize eps
    return {
        "means3D": raw["xyz"].astype(np.float32),
        "scales": np.exp(raw["scaling"].astype(np.float32)),
        "rotations": (q / n).astype(np.float32),
        "opacity": (1.0 / (1.0 + np.exp(-raw["opacity"].astype(np.float64)))).astype(np.float32),
        "shs": np.ascontiguousarray(np.concatenate([raw["features_dc"], raw["features_rest"]], axis=1).astype(np.float32)),
        "sh_degree": int(raw["sh_degree"]),
    }


def model_from_ply(path, max_sh_degree=3):
    """point_cloud.ply -> rasterizer inputs (raw parameters + activations)."""
    return activate(raw_model_from_ply(path, max_sh_degree))


def compose_levels(level_models):
    """compose_models.py:39-80.  `level_models[0]` is the finest (PS1) model, `level_models[i]` (i >= 1) the level-i model
    with `indexes` into level 0 (dicts from raw_model_from_ply).  Returns the foveated scene dict the `fov` rasterizer takes:
    geometry and SH-rest of level 0, highest_levels [P,1] float32, shs_dcs [P,L,3] (RAW dc coefficient per level),
    opacities4 [P,L] (activated), shs_rest [P,15,3]."""
    L = len(level_models)
    base = activate(level_models[0])
    P = base["means3D"].shape[0]
    shs_dcs = np.zeros((P, L, 3), np.float32)
    highest = np.zeros((P, 1), np.float32)
    opac = np.ones((P, L), np.float32)
    shs_dcs[:, 0, :] = base["shs"][:, 0, :]
    opac[:, 0] = base["opacity"][:, 0]
    for i in range(1, L):
        m = activate(level_models[i])
        idx = level_models[i]["indexes"].reshape(-1).astype(np.int64)
        shs_dcs[:, i, :] = shs_dcs[:, i - 1, :]
        shs_dcs[idx, i, :] = m["shs"][:, 0, :]
        opac[:, i] = opac[:, i - 1]
        opac[idx, i] = m["opacity"][:, 0]
        highest[idx] = i
    out = dict(base)
    out["highest_levels"] = highest
    out["shs_dcs"] = shs_dcs
    out["opacities4"] = opac
    out["shs_rest"] = np.ascontiguousarray(base["shs"][:, 1:, :])
    return out


def save_composed(folder, composed):
    """Writes highest_levels.pt / shs_dcs.pt / opacities.pt as compose_models.py:75-80 does (CPU float32 torch tensors), so that
    render_compose_gazes_fps.py:85-96 (`torch.load(...).cuda()`) reads them unchanged."""
    import torch
    os.makedirs(folder, exist_ok=True)
    for name, key in (("highest_levels.pt", "highest_levels"), ("shs_dcs.pt", "shs_dcs"), ("opacities.pt", "opacities4")):
        torch.save(torch.from_numpy(np.ascontiguousarray(composed[key], dtype=np.float32)), os.path.join(folder, name))


def load_composed(folder):
    """The inverse: {"highest_levels" [P,1], "shs_dcs" [P,L,3], "opacities4" [P,L]} as float32 numpy arrays."""
    import torch
    out = {}
    for name, key in (("highest_levels.pt", "highest_levels"), ("shs_dcs.pt", "shs_dcs"), ("opacities.pt", "opacities4")):
        t = torch.load(os.path.join(folder, name), map_location="cpu")
        out[key] = np.ascontiguousarray(t.detach().float().numpy())
    P = out["shs_dcs"].shape[0]
    if out["highest_levels"].reshape(-1).shape[0] != P or out["opacities4"].shape != out["shs_dcs"].shape[:2]:
        raise ValueError(f"{folder}: highest_levels / shs_dcs / opacities do not describe the same model")
    out["highest_levels"] = out["highest_levels"].reshape(P, 1)
    return out


def load_foveated_model(base_folder, layer_num, max_pooling_size, iteration=55000, max_sh_degree=3):
    """The model files render_compose_gazes_fps.py:80-90 reads, as the scene dict `forward_fov` takes:
    `<base>/1_PS1_<L>_<S>/point_cloud/iteration_<it>/point_cloud.ply` (geometry, SH rest) and
    `<base>/composed_<L>_<S>/{highest_levels,shs_dcs,opacities}.pt`."""
    ply = os.path.join(base_folder, f"1_PS1_{layer_num}_{max_pooling_size}", "point_cloud", f"iteration_{iteration}", "point_cloud.ply")
    m = model_from_ply(ply, max_sh_degree)
    c = load_composed(os.path.join(base_folder, f"composed_{layer_num}_{max_pooling_size}"))
    P = m["means3D"].shape[0]
    if c["shs_dcs"].shape[0] != P:
        raise ValueError(f"composed tensors describe {c['shs_dcs'].shape[0]} Gaussians, {ply} has {P}")
    out = {"means3D": m["means3D"], "scales": m["scales"], "rotations": m["rotations"], "sh_degree": m["sh_degree"],
           "shs_rest": np.ascontiguousarray(m["shs"][:, 1:, :]), "opacity": m["opacity"], "shs": m["shs"]}
    out.update(c)
    return out


def smfr_levels(level_sizes, seed=None):
    """gen_naive_FR.py:33-59.  `level_sizes[0]` = P of the shared model, `level_sizes[i]` = number of Gaussians level i keeps.
    Level i's subset is the first `level_sizes[i]` entries of level i-1's (randomly permuted) subset, so the subsets nest;
    returns highest_levels [P] float32 (the reference saves a 1-D tensor as highest_levels.pth)."""
    P = int(level_sizes[0])
    rng = np.random.default_rng(seed)
    current = rng.permutation(P)
    highest = np.zeros(P, np.float32)
    for i in range(1, len(level_sizes)):
        n = int(level_sizes[i])
        if n > len(current):
            raise ValueError(f"level {i} keeps {n} Gaussians but level {i - 1} has only {len(current)}")
        current = current[:n]
        highest[current] = i
    return highest


def _projection(znear, zfar, fovx, fovy):
    """utils/graphics_utils.py:51-71 getProjectionMatrix (float32 like the torch original)."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2.0 * znear / (right + right)
    P[1, 1] = 2.0 * znear / (top + top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_from_rt(R, T, fovx, fovy, width, height, znear=0.01, zfar=100.0):
    """scene/cameras.py:17-57 from (R, T, FoVx, FoVy): R is camera-to-world (columns = camera axes), T the world-to-camera
    translation, both as the COLMAP readers produce them."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = np.asarray(R, np.float64).T
    Rt[:3, 3] = np.asarray(T, np.float64)
    Rt[3, 3] = 1.0
    wv = np.float32(Rt).T.copy()                                   # world_view_transform (stored transposed)
    pj = _projection(znear, zfar, fovx, fovy).T.copy()
    full = (wv @ pj).astype(np.float32)                            # full_proj_transform
    campos = np.linalg.inv(wv.astype(np.float64))[3, :3].astype(np.float32)
    return {
        "image_width": int(width), "image_height": int(height), "tanfovx": math.tan(fovx * 0.5), "tanfovy": math.tan(fovy * 0.5),
        "viewmatrix": np.ascontiguousarray(wv), "projmatrix": np.ascontiguousarray(full), "campos": np.ascontiguousarray(campos),
        "FoVx": float(fovx), "FoVy": float(fovy),
    }


def camera_from_json_entry(e, znear=0.01, zfar=100.0):
    """One cameras.json entry (camera_to_JSON, utils/camera_utils.py:62-82): `rotation` / `position` are the camera-to-world
    rotation and the camera centre, fx / fy focal lengths in pixels."""
    R = np.asarray(e["rotation"], np.float64)                      # = Camera.R
    pos = np.asarray(e["position"], np.float64)
    T = -R.T @ pos
    W, H = int(e["width"]), int(e["height"])
    fovx = 2.0 * math.atan(W / (2.0 * float(e["fx"])))             # focal2fov
    fovy = 2.0 * math.atan(H / (2.0 * float(e["fy"])))
    cam = camera_from_rt(R, T, fovx, fovy, W, H, znear, zfar)
    cam["id"] = e.get("id")
    cam["img_name"] = e.get("img_name")
    return cam


def cameras_from_json(path):
    with open(path) as f:
        return [camera_from_json_entry(e) for e in json.load(f)]


def camera_to_json_entry(cam_id, R, T, fovx, fovy, width, height, img_name=""):
    """camera_to_JSON restated (for writing fixtures / round trips)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = np.asarray(R, np.float64).T
    Rt[:3, 3] = np.asarray(T, np.float64)
    Rt[3, 3] = 1.0
    c2w = np.linalg.inv(Rt)
    return {"id": cam_id, "img_name": img_name, "width": int(width), "height": int(height), "position": c2w[:3, 3].tolist(),
            "rotation": [r.tolist() for r in c2w[:3, :3]], "fy": height / (2 * math.tan(fovy / 2)), "fx": width / (2 * math.tan(fovx / 2))}
