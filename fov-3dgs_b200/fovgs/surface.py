"""The reference's Python operator surface, re-stated once and instantiated per package name.

Each drop-in package (diff_gaussian_rasterization_fov_pcheck_obb, ..._pcheck_obb, ..._pcheck_obb_sum) exports
exactly what the reference package exports — GaussianRasterizationSettings, GaussianRasterizer,
rasterize_gaussians, _RasterizeGaussians — with the same argument names, positional order, return arity and error
behaviour (reference: FOV/diff_gaussian_rasterization_fov_pcheck_obb/__init__.py:21-261,
SUM/diff_gaussian_rasterization_pcheck_obb_sum/__init__.py:26-226), so fov3dgs/gaussian_renderer*/__init__.py and
render_compose_gazes_fps*.py run unchanged with `fov-3dgs_b200/` first on sys.path.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import ops


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def _check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp):
    if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')


_EMPTY = torch.Tensor([])


def _empty_if_none(t):
    # the reference builds a fresh torch.Tensor([]) per missing argument (~2 us each); one shared, never-written instance does
    return _EMPTY if t is None else t


class _NoGradCtx:
    """Stands in for the autograd context while gradients are disabled (render_compose_gazes_fps*.py, render.py and prune.py
    run under torch.no_grad()): Function.apply then only adds per-argument bookkeeping (~25 us per frame) to a forward whose
    outputs carry no graph either way."""

    def mark_non_differentiable(self, *tensors):
        pass

    def save_for_backward(self, *tensors):
        pass


def _apply(fn, *args):
    if not torch.is_grad_enabled():
        return fn.forward(_NoGradCtx(), *args)
    return fn.apply(*args)


def _mark_visible(raster_settings, positions):
    with torch.no_grad():
        return ops.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)


# ------------------------------------------------------------------------------------------------------------
# foveated, forward-only (configs 3/4)
# ------------------------------------------------------------------------------------------------------------
class _RasterizeGaussiansFov(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs_rest, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, shs_dcs, highest_levels, gazeArray, alpha, blending):
        if colors_precomp is not None and colors_precomp.numel() != 0:
            raise RuntimeError("the foveated rasterizer renders from SH coefficients; colors_precomp is not supported "
                               "(the reference ignores it: FOV/cuda_rasterizer/forward.cu:214-222)")
        if cov3Ds_precomp is not None and cov3Ds_precomp.numel() != 0:
            raise RuntimeError("the foveated rasterizer takes scales/rotations, not cov3D_precomp")
        args = (shs_dcs, highest_levels, gazeArray, alpha, blending, raster_settings.bg, means3D, colors_precomp,
                opacities, scales, rotations, raster_settings.scale_modifier, cov3Ds_precomp,
                raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width, shs_rest,
                raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)

        def run():
            return ops.forward_fov(means3D, opacities, scales, rotations, shs_rest, shs_dcs, highest_levels, gazeArray,
                                   alpha, blending, raster_settings)

        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                num_rendered, color, radii = run()
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, radii = run()
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        # forward-only variant: the reference stubs its backward with Nones (FOV/.../__init__.py:128-187, Q8)
        return (None,) * 14


def _fov_rasterize_gaussians(means3D, means2D, shs_rest, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                             raster_settings, shs_dcs, highest_levels, gazeArray, alpha: float, blending: bool):
    return _apply(_RasterizeGaussiansFov, means3D, means2D, shs_rest, colors_precomp, opacities, scales, rotations,
                                        cov3Ds_precomp, raster_settings, shs_dcs, highest_levels, gazeArray, alpha,
                                        blending)


class _FovGaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        object.__setattr__(self, "raster_settings", raster_settings)   # a plain attribute: nn.Module.__setattr__ costs ~6 us per frame

    def markVisible(self, positions):
        return _mark_visible(self.raster_settings, positions)

    # Extension (not in the reference): `rasterizer.output_uint8 = True` makes forward() return the uint8 image the reference's
    # scripts store (torchvision.utils.save_image's quantisation of the fp32 render, fov3dgs/render.py:52), written by the blend
    # epilogue in place of the fp32 image — 6.2 MB instead of 24.9 MB per 1080p frame on its way to the host.  Inference only.
    output_uint8 = False

    def forward(self, means3D, means2D, opacities, shs_rest=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, shs_dcs=None, highest_levels=None, gazeArray=None, alpha=None, blending=None):
        _check_inputs(shs_rest, colors_precomp, scales, rotations, cov3D_precomp)
        if self.output_uint8:
            if colors_precomp is not None or cov3D_precomp is not None:
                raise RuntimeError("the foveated rasterizer renders from SH coefficients and scales/rotations")
            _, color, radii = ops.forward_fov(means3D, opacities, scales, rotations, shs_rest, shs_dcs, highest_levels, gazeArray,
                                              alpha, blending, self.raster_settings, out_uint8=True)
            return color, radii
        return _fov_rasterize_gaussians(means3D, means2D, _empty_if_none(shs_rest), _empty_if_none(colors_precomp),
                                        opacities, _empty_if_none(scales), _empty_if_none(rotations),
                                        _empty_if_none(cov3D_precomp), self.raster_settings, shs_dcs, highest_levels,
                                        gazeArray, alpha, blending)


def make_fov_api():
    return {
        "GaussianRasterizationSettings": GaussianRasterizationSettings,
        "GaussianRasterizer": _FovGaussianRasterizer,
        "rasterize_gaussians": _fov_rasterize_gaussians,
        "_RasterizeGaussians": _RasterizeGaussiansFov,
        "cpu_deep_copy_tuple": cpu_deep_copy_tuple,
    }


# ------------------------------------------------------------------------------------------------------------
# SMFR baseline: foveated forward with one shared model (diff_gaussian_rasterization_naive_pcheck_obb)
# ------------------------------------------------------------------------------------------------------------
class _RasterizeGaussiansSmfr(torch.autograd.Function):
    """naive_pcheck_obb/diff_gaussian_rasterization_naive_pcheck_obb/__init__.py:56-131 (forward-only like the FOV variant)."""

    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
                highest_levels, gazeArray, alpha, blending):
        if colors_precomp is not None and colors_precomp.numel() != 0:
            raise RuntimeError("the shared-model foveated rasterizer renders from SH coefficients; colors_precomp is not supported")
        if cov3Ds_precomp is not None and cov3Ds_precomp.numel() != 0:
            raise RuntimeError("the shared-model foveated rasterizer takes scales/rotations, not cov3D_precomp")
        args = (highest_levels, gazeArray, alpha, blending, raster_settings.bg, means3D, colors_precomp, opacities, scales,
                rotations, raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy, raster_settings.image_height,
                raster_settings.image_width, shs, raster_settings.sh_degree, raster_settings.campos,
                raster_settings.prefiltered, raster_settings.debug)

        def run():
            return ops.forward_smfr(means3D, opacities, scales, rotations, shs, highest_levels, gazeArray, alpha, blending,
                                    raster_settings)

        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                num_rendered, color, radii = run()
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, radii = run()
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        return (None,) * 13


def _smfr_rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                              raster_settings, highest_levels, gazeArray, alpha: float, blending: bool):
    return _apply(_RasterizeGaussiansSmfr, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                         raster_settings, highest_levels, gazeArray, alpha, blending)


class _SmfrGaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        object.__setattr__(self, "raster_settings", raster_settings)   # a plain attribute: nn.Module.__setattr__ costs ~6 us per frame

    def markVisible(self, positions):
        return _mark_visible(self.raster_settings, positions)

    # `shs_dcs` is accepted and unused, exactly like the reference signature (…/__init__.py:214-216)
    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, shs_dcs=None, highest_levels=None, gazeArray=None, alpha=None, blending=None):
        _check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
        return _smfr_rasterize_gaussians(means3D, means2D, _empty_if_none(shs), _empty_if_none(colors_precomp), opacities,
                                         _empty_if_none(scales), _empty_if_none(rotations), _empty_if_none(cov3D_precomp),
                                         self.raster_settings, highest_levels, gazeArray, alpha, blending)


def make_smfr_api():
    return {
        "GaussianRasterizationSettings": GaussianRasterizationSettings,
        "GaussianRasterizer": _SmfrGaussianRasterizer,
        "rasterize_gaussians": _smfr_rasterize_gaussians,
        "_RasterizeGaussians": _RasterizeGaussiansSmfr,
        "cpu_deep_copy_tuple": cpu_deep_copy_tuple,
    }


# ------------------------------------------------------------------------------------------------------------
# MMFR baseline: one call per level model (diff_gaussian_rasterization_mmfr_pcheck_obb)
# ------------------------------------------------------------------------------------------------------------
class _RasterizeGaussiansMmfr(torch.autograd.Function):
    """mmfr_pcheck_obb/diff_gaussian_rasterization_mmfr_pcheck_obb/__init__.py:56-131 (forward-only)."""

    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
                cur_level, gazeArray, alpha, blending):
        if colors_precomp is not None and colors_precomp.numel() != 0:
            raise RuntimeError("the multi-model foveated rasterizer renders from SH coefficients; colors_precomp is not supported")
        if cov3Ds_precomp is not None and cov3Ds_precomp.numel() != 0:
            raise RuntimeError("the multi-model foveated rasterizer takes scales/rotations, not cov3D_precomp")
        args = (cur_level, gazeArray, alpha, blending, raster_settings.bg, means3D, colors_precomp, opacities, scales,
                rotations, raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy, raster_settings.image_height,
                raster_settings.image_width, shs, raster_settings.sh_degree, raster_settings.campos,
                raster_settings.prefiltered, raster_settings.debug)

        def run():
            return ops.forward_mmfr(means3D, opacities, scales, rotations, shs, cur_level, gazeArray, alpha, blending,
                                    raster_settings)

        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                num_rendered, color, radii = run()
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, radii = run()
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        return (None,) * 13


def _mmfr_rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                              raster_settings, cur_level, gazeArray, alpha: float, blending: bool):
    return _apply(_RasterizeGaussiansMmfr, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                         raster_settings, cur_level, gazeArray, alpha, blending)


class _MmfrGaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        object.__setattr__(self, "raster_settings", raster_settings)   # a plain attribute: nn.Module.__setattr__ costs ~6 us per frame

    def markVisible(self, positions):
        return _mark_visible(self.raster_settings, positions)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, shs_dcs=None, cur_level=None, gazeArray=None, alpha=None, blending=None):
        _check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
        return _mmfr_rasterize_gaussians(means3D, means2D, _empty_if_none(shs), _empty_if_none(colors_precomp), opacities,
                                         _empty_if_none(scales), _empty_if_none(rotations), _empty_if_none(cov3D_precomp),
                                         self.raster_settings, cur_level, gazeArray, alpha, blending)


def make_mmfr_api():
    return {
        "GaussianRasterizationSettings": GaussianRasterizationSettings,
        "GaussianRasterizer": _MmfrGaussianRasterizer,
        "rasterize_gaussians": _mmfr_rasterize_gaussians,
        "_RasterizeGaussians": _RasterizeGaussiansMmfr,
        "cpu_deep_copy_tuple": cpu_deep_copy_tuple,
    }


# ------------------------------------------------------------------------------------------------------------
# PS=1: inference (pcheck_obb) and training (pcheck_obb_sum)
# ------------------------------------------------------------------------------------------------------------
def _make_ps1_function(mode: int):
    """mode: ops.MODE_OBB (inference), one of the training family ops.MODE_SUM / MODE_MAX / MODE_LWMC, or ops.MODE_VANILLA
    (the stock diff_gaussian_rasterization: forward + backward, returns (color, radii) like the reference's
    fov3dgs/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py:83-99).  The
    loss-weighted variant takes one more argument, `loss_map`, after `raster_settings`
    (.../pcheck_obb_loss_weighted_max_count/diff_gaussian_rasterization_pcheck_obb_loss_weighted_max_count/__init__.py:24-47)."""
    sum_mode = mode != ops.MODE_OBB          # keeps the state a backward needs
    stat_mode = sum_mode and mode != ops.MODE_VANILLA   # returns gaussians_count / contributions as well
    with_loss_map = mode == ops.MODE_LWMC

    class _RasterizeGaussians(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    raster_settings, *extra):
            loss_map = extra[0] if with_loss_map else None
            args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                    raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                    raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                    raster_settings.image_height, raster_settings.image_width, sh, raster_settings.sh_degree,
                    raster_settings.campos, raster_settings.prefiltered) + \
                   ((loss_map,) if with_loss_map else ()) + (raster_settings.debug,)

            def run():
                return ops.forward_ps1(mode, means3D, opacities, scales, rotations, cov3Ds_precomp, sh, colors_precomp,
                                       raster_settings, loss_map=loss_map)

            if raster_settings.debug:
                cpu_args = cpu_deep_copy_tuple(args)
                try:
                    out = run()
                except Exception as ex:
                    torch.save(cpu_args, "snapshot_fw.dump")
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                    raise ex
            else:
                out = run()
            num_rendered, color, radii, item = out[:4]
            ctx.raster_settings = raster_settings
            ctx.num_rendered = num_rendered
            ctx.workspace_item = item  # opaque saved state (replaces geomBuffer/binningBuffer/imgBuffer)
            ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh)
            if stat_mode:
                gaussians_count, contributions = out[4], out[5]
                ctx.mark_non_differentiable(radii, gaussians_count, contributions)
                return color, radii, gaussians_count, contributions
            ctx.mark_non_differentiable(radii)
            return color, radii

        @staticmethod
        def backward(ctx, grad_out_color, *_unused):
            if not sum_mode:
                raise RuntimeError(
                    "diff_gaussian_rasterization_pcheck_obb is the inference variant: its forward does not keep "
                    "final_T / n_contrib (reference OBB/cuda_rasterizer/forward.cu:379-380), so no gradient exists. "
                    "Use diff_gaussian_rasterization_pcheck_obb_sum for training.")
            raster_settings = ctx.raster_settings
            colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh = ctx.saved_tensors

            def run():
                return ops.backward_ps1(ctx.workspace_item, means3D, radii, scales, rotations, cov3Ds_precomp, sh,
                                        colors_precomp, raster_settings, grad_out_color)

            if raster_settings.debug:
                cpu_args = cpu_deep_copy_tuple((raster_settings.bg, means3D, radii, colors_precomp, scales, rotations,
                                                raster_settings.scale_modifier, cov3Ds_precomp,
                                                raster_settings.viewmatrix, raster_settings.projmatrix,
                                                raster_settings.tanfovx, raster_settings.tanfovy, grad_out_color, sh,
                                                raster_settings.sh_degree, raster_settings.campos, ctx.num_rendered,
                                                raster_settings.debug))
                try:
                    res = run()
                except Exception as ex:
                    torch.save(cpu_args, "snapshot_bw.dump")
                    print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                    raise ex
            else:
                res = run()
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
             grad_rotations) = res
            return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                    grad_rotations, grad_cov3Ds_precomp, None) + ((None,) if with_loss_map else ())

    return _RasterizeGaussians


def _make_ps1_api(mode: int):
    Fn = _make_ps1_function(mode)

    if mode == ops.MODE_LWMC:
        def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                raster_settings, loss_map):
            return _apply(Fn, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                            raster_settings, loss_map)

        class GaussianRasterizer(nn.Module):
            def __init__(self, raster_settings):
                super().__init__()
                object.__setattr__(self, "raster_settings", raster_settings)   # a plain attribute: nn.Module.__setattr__ costs ~6 us per frame

            def markVisible(self, positions):
                return _mark_visible(self.raster_settings, positions)

            def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                        cov3D_precomp=None, loss_map=None):
                _check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
                return rasterize_gaussians(means3D, means2D, _empty_if_none(shs), _empty_if_none(colors_precomp),
                                           opacities, _empty_if_none(scales), _empty_if_none(rotations),
                                           _empty_if_none(cov3D_precomp), self.raster_settings, loss_map)

        return {
            "GaussianRasterizationSettings": GaussianRasterizationSettings,
            "GaussianRasterizer": GaussianRasterizer,
            "rasterize_gaussians": rasterize_gaussians,
            "_RasterizeGaussians": Fn,
            "cpu_deep_copy_tuple": cpu_deep_copy_tuple,
        }

    def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                            raster_settings):
        return _apply(Fn, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings)

    class GaussianRasterizer(nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            object.__setattr__(self, "raster_settings", raster_settings)   # a plain attribute: nn.Module.__setattr__ costs ~6 us per frame

        def markVisible(self, positions):
            return _mark_visible(self.raster_settings, positions)

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            _check_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
            return rasterize_gaussians(means3D, means2D, _empty_if_none(shs), _empty_if_none(colors_precomp), opacities,
                                       _empty_if_none(scales), _empty_if_none(rotations), _empty_if_none(cov3D_precomp),
                                       self.raster_settings)

    return {
        "GaussianRasterizationSettings": GaussianRasterizationSettings,
        "GaussianRasterizer": GaussianRasterizer,
        "rasterize_gaussians": rasterize_gaussians,
        "_RasterizeGaussians": Fn,
        "cpu_deep_copy_tuple": cpu_deep_copy_tuple,
    }


def make_obb_api():
    return _make_ps1_api(ops.MODE_OBB)


def make_sum_api():
    return _make_ps1_api(ops.MODE_SUM)


def make_max_api():
    return _make_ps1_api(ops.MODE_MAX)


def make_lwmc_api():
    return _make_ps1_api(ops.MODE_LWMC)


def make_vanilla_api():
    return _make_ps1_api(ops.MODE_VANILLA)
