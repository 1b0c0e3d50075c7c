"""The drop-in packages expose the reference's operator surface: same names, argument order, error behaviour
(reference: FOV/.../__init__.py:21-261, SUM/.../__init__.py:26-226, fov3dgs/gaussian_wrapper.py:2-7)."""
import importlib
import inspect

import pytest
import torch

PKGS = ["diff_gaussian_rasterization_fov_pcheck_obb", "diff_gaussian_rasterization_naive_pcheck_obb",
        "diff_gaussian_rasterization_mmfr_pcheck_obb",
        "diff_gaussian_rasterization_pcheck_obb",
        "diff_gaussian_rasterization_pcheck_obb_sum", "diff_gaussian_rasterization_pcheck_obb_max",
        "diff_gaussian_rasterization_pcheck_obb_loss_weighted_max_count", "diff_gaussian_rasterization"]


@pytest.mark.parametrize("name", PKGS)
def test_packages_import_and_export_the_three_names(name):
    m = importlib.import_module(name)
    for attr in ("GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"):
        assert hasattr(m, attr)
    fields = m.GaussianRasterizationSettings._fields
    assert fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                      "projmatrix", "sh_degree", "campos", "prefiltered", "debug")


def test_fov_signatures_match_reference():
    m = importlib.import_module("diff_gaussian_rasterization_fov_pcheck_obb")
    assert list(inspect.signature(m.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "shs_rest", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings", "shs_dcs", "highest_levels", "gazeArray", "alpha", "blending"]
    assert list(inspect.signature(m.GaussianRasterizer.forward).parameters) == [
        "self", "means3D", "means2D", "opacities", "shs_rest", "colors_precomp", "scales", "rotations", "cov3D_precomp",
        "shs_dcs", "highest_levels", "gazeArray", "alpha", "blending"]


def test_smfr_signatures_match_reference():
    """naive_pcheck_obb/diff_gaussian_rasterization_naive_pcheck_obb/__init__.py:20-36,214-216."""
    m = importlib.import_module("diff_gaussian_rasterization_naive_pcheck_obb")
    assert list(inspect.signature(m.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings", "highest_levels", "gazeArray", "alpha", "blending"]
    assert list(inspect.signature(m.GaussianRasterizer.forward).parameters) == [
        "self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp",
        "shs_dcs", "highest_levels", "gazeArray", "alpha", "blending"]


def test_mmfr_signatures_match_reference():
    """mmfr_pcheck_obb/diff_gaussian_rasterization_mmfr_pcheck_obb/__init__.py:20-36,215-217."""
    m = importlib.import_module("diff_gaussian_rasterization_mmfr_pcheck_obb")
    assert list(inspect.signature(m.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "shs", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings", "cur_level", "gazeArray", "alpha", "blending"]
    assert list(inspect.signature(m.GaussianRasterizer.forward).parameters) == [
        "self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp",
        "shs_dcs", "cur_level", "gazeArray", "alpha", "blending"]


@pytest.mark.parametrize("name", ["diff_gaussian_rasterization_pcheck_obb", "diff_gaussian_rasterization_pcheck_obb_sum",
                                  "diff_gaussian_rasterization_pcheck_obb_max"])
def test_ps1_signatures_match_reference(name):
    m = importlib.import_module(name)
    assert list(inspect.signature(m.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp", "raster_settings"]
    assert list(inspect.signature(m.GaussianRasterizer.forward).parameters) == [
        "self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp"]
    assert hasattr(m.GaussianRasterizer, "markVisible")


def test_loss_weighted_signature_matches_reference():
    """.../pcheck_obb_loss_weighted_max_count/.../__init__.py:24-36,198: `loss_map` is the last argument."""
    m = importlib.import_module("diff_gaussian_rasterization_pcheck_obb_loss_weighted_max_count")
    assert list(inspect.signature(m.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp", "raster_settings",
        "loss_map"]
    assert list(inspect.signature(m.GaussianRasterizer.forward).parameters) == [
        "self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp", "loss_map"]


def _settings(m):
    z = torch.zeros(3)
    return m.GaussianRasterizationSettings(16, 16, 1.0, 1.0, z, 1.0, torch.eye(4), torch.eye(4), 3, z, False, False)


@pytest.mark.parametrize("name", ["diff_gaussian_rasterization_fov_pcheck_obb", "diff_gaussian_rasterization_pcheck_obb_sum"])
def test_argument_validation_raises_like_reference(name):
    m = importlib.import_module(name)
    r = m.GaussianRasterizer(raster_settings=_settings(m))
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1))
    kw = {"shs_rest": torch.zeros(4, 15, 3)} if "fov" in name else {"shs": torch.zeros(4, 16, 3)}
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), **kw)


def test_cpu_tensors_are_rejected_not_silently_computed():
    """There is no CPU fallback on the product path: CPU inputs raise."""
    m = importlib.import_module("diff_gaussian_rasterization_pcheck_obb")
    r = m.GaussianRasterizer(raster_settings=_settings(m))
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), scales=x, rotations=torch.zeros(4, 4))


def test_vanilla_package_is_a_real_rasterizer():
    """fov3dgs/gaussian_wrapper.py:2,11: cuda_type="original" constructs diff_gaussian_rasterization.GaussianRasterizer — same
    surface as the reference's stock package (forward signature, CPU tensors rejected: no fallback)."""
    m = importlib.import_module("diff_gaussian_rasterization")
    r = m.GaussianRasterizer(raster_settings=_settings(m))
    assert list(inspect.signature(r.forward).parameters) == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                                             "rotations", "cov3D_precomp"]
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        r(means3D=x, means2D=x, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), scales=x, rotations=torch.zeros(4, 4))


def test_gaussian_wrapper_import_line_works():
    """fov3dgs/gaussian_wrapper.py:2-7 imports six names at module import; all must resolve."""
    for n in PKGS[3:]:
        importlib.import_module(n).GaussianRasterizer
