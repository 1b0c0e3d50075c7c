"""fovgs — B200-native (sm_100a) foveated 3D Gaussian Splatting rasterizer behind the reference operator surface.

`fovgs.ops`      torch-facing operators over the C-ABI (include/fovgs.h, lib/libfovgs.so)
`fovgs.surface`  GaussianRasterizationSettings / GaussianRasterizer / rasterize_gaussians factories
The sibling packages diff_gaussian_rasterization_* are the drop-in names the reference scripts import.
"""
from . import _lib  # noqa: F401
from . import ops, surface  # noqa: F401

__all__ = ["ops", "surface"]
