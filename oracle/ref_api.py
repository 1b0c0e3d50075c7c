"""Test infrastructure: thin callers/decoders for the UNMODIFIED reference `_C` modules built by build_ref.py.

Only tests/, tools/ and bench.py --impl reference import this.  Nothing here is on the product path.

Argument orders are the reference's pybind signatures:
  FOV  rasterize_gaussians(24 args)  FOV/rasterize_points.h:17-44
  PS1  rasterize_gaussians(19 args)  SUM/rasterize_points.cu:35-55 ; backward(21 args) SUM/rasterize_points.cu:137-159
Buffer decoding follows GeometryState/ImageState/BinningState::fromChunk
  (FOV/cuda_rasterizer/rasterizer_impl.cu:574-615, SUM/cuda_rasterizer/rasterizer_impl.cu:246-286).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import build_ref  # noqa: E402

_mods = {}


def ref_module(name):
    if name not in _mods:
        _mods[name] = build_ref.load_ref(name)
    return _mods[name]


def available(name):
    return build_ref.so_path(name) is not None


def _empty():
    return torch.Tensor([])


def fov_forward(mod, sc, cam, gaze, alpha=0.05, blending=True, bg=None, debug=False):
    """sc: dict of CUDA tensors (means3D, scales, rotations, opacities4, shs_rest, shs_dcs, highest_levels)."""
    bg = bg if bg is not None else torch.zeros(3, device="cuda")
    return mod.rasterize_gaussians(
        sc["shs_dcs"], sc["highest_levels"], gaze, float(alpha), bool(blending), bg, sc["means3D"], _empty(),
        sc["opacities4"], sc["scales"], sc["rotations"], 1.0, _empty(), cam["viewmatrix"], cam["projmatrix"],
        cam["tanfovx"], cam["tanfovy"], cam["image_height"], cam["image_width"], sc["shs_rest"], sc["sh_degree"],
        cam["campos"], False, debug)


def smfr_forward(mod, sc, cam, gaze, alpha=0.05, blending=True, bg=None, debug=False):
    """ref_naive_C (SMFR baseline): 23 args, naive_pcheck_obb/rasterize_points.h:17-43.  sc carries shs [P,M,3],
    opacity [P,1], highest_levels [P,1]."""
    bg = bg if bg is not None else torch.zeros(3, device="cuda")
    return mod.rasterize_gaussians(
        sc["highest_levels"], gaze, float(alpha), bool(blending), bg, sc["means3D"], _empty(), sc["opacity"], sc["scales"],
        sc["rotations"], 1.0, _empty(), cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"],
        cam["image_height"], cam["image_width"], sc["shs"], sc["sh_degree"], cam["campos"], False, debug)


def mmfr_forward(mod, sc, cam, cur_level, gaze, alpha=0.05, blending=True, bg=None, debug=False):
    """ref_mmfr_C (MMFR baseline, one level call): 23 args, mmfr_pcheck_obb/rasterize_points.h:17-43.  NOTE: the reference
    keeps its tile tables in process-static memory and refreshes them only when cur_level == 0 — call level 0 first."""
    bg = bg if bg is not None else torch.zeros(3, device="cuda")
    return mod.rasterize_gaussians(
        float(cur_level), gaze, float(alpha), bool(blending), bg, sc["means3D"], _empty(), sc["opacity"], sc["scales"],
        sc["rotations"], 1.0, _empty(), cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"],
        cam["image_height"], cam["image_width"], sc["shs"], sc["sh_degree"], cam["campos"], False, debug)


def ps1_forward(mod, sc, cam, bg=None, debug=False, loss_map=None):
    """`loss_map` (CUDA [H,W]) only for ref_lwmc_C, whose pybind signature takes it between `prefiltered` and `debug`
    (.../pcheck_obb_loss_weighted_max_count/rasterize_points.cu:35-57)."""
    bg = bg if bg is not None else torch.zeros(3, device="cuda")
    tail = (False, debug) if loss_map is None else (False, loss_map, debug)
    return mod.rasterize_gaussians(
        bg, sc["means3D"], _empty(), sc["opacity"], sc["scales"], sc["rotations"], 1.0, _empty(), cam["viewmatrix"],
        cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], cam["image_height"], cam["image_width"], sc["shs"],
        sc["sh_degree"], cam["campos"], *tail)


def ps1_backward(mod, sc, cam, radii, grad_out, geom, num_rendered, binning, img, bg=None, debug=False):
    bg = bg if bg is not None else torch.zeros(3, device="cuda")
    return mod.rasterize_gaussians_backward(
        bg, sc["means3D"], radii, _empty(), sc["scales"], sc["rotations"], 1.0, _empty(), cam["viewmatrix"],
        cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], grad_out, sc["shs"], sc["sh_degree"], cam["campos"], geom,
        num_rendered, binning, img, debug)


class _Chunk:
    def __init__(self, buf):
        self.buf = buf
        self.base = buf.data_ptr()
        self.off = 0

    def take(self, count, dtype, elems=1):
        itemsize = torch.tensor([], dtype=dtype).element_size() * elems
        addr = (self.base + self.off + 127) & ~127
        start = addr - self.base
        nbytes = count * itemsize
        out = self.buf[start:start + nbytes].view(dtype)
        self.off = start + nbytes
        return out.view(count, elems) if elems > 1 else out

    def skip(self, nbytes):
        addr = (self.base + self.off + 127) & ~127
        self.off = addr - self.base + nbytes


def decode_geom(geom, P, variant, scan_bytes=None):
    """variant: 'fov' | 'ps1'.  Returns dict of views (depths, means2D, cov3D, conic, rgb, tiles_touched)."""
    c = _Chunk(geom)
    out = {}
    out["depths"] = c.take(P, torch.float32)
    out["clamped"] = c.take(3 * P, torch.uint8)
    out["internal_radii"] = c.take(P, torch.int32)
    out["means2D"] = c.take(P, torch.float32, 2)
    out["cov3D"] = c.take(P, torch.float32, 6)
    if variant == "fov":
        out["conic"] = c.take(P, torch.float32, 3)
    else:
        co = c.take(P, torch.float32, 4)
        out["conic"] = co[:, :3]
        out["conic_opacity"] = co
    out["rgb"] = c.take(P, torch.float32, 3)
    out["tiles_touched"] = c.take(P, torch.int32)
    return out


def decode_binning(binning, N):
    c = _Chunk(binning)
    out = {}
    out["point_list"] = c.take(N, torch.int32)
    out["point_list_unsorted"] = c.take(N, torch.int32)
    out["keys"] = c.take(N, torch.int64)
    out["keys_unsorted"] = c.take(N, torch.int64)
    return out


def decode_img(img, W, H):
    c = _Chunk(img)
    n = W * H
    out = {}
    out["accum_alpha"] = c.take(n, torch.float32)
    out["n_contrib"] = c.take(n, torch.int32)
    out["ranges"] = c.take(n, torch.int32, 2)
    return out
