"""Drop-in `diff_gaussian_rasterization` for nyu-systems/Grendel-GS, backed by the B200-native C-ABI library.

Exports exactly the names the reference imports:
  gaussian_renderer/__init__.py:14-17      GaussianRasterizationSettings, GaussianRasterizer
  gaussian_renderer/workload_division.py:6 diff_gaussian_rasterization._C.get_local2j_ids_bool(...)
  arguments/__init__.py:17,254-257         diff_gaussian_rasterization._C.get_block_XY()
Put `<repo>/grendel-gs_b200` on PYTHONPATH and the reference's train.py imports this package unchanged.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from gs_b200 import ops as _ops
from . import _C  # noqa: F401  (reference code reaches it as diff_gaussian_rasterization._C)


class GaussianRasterizationSettings(NamedTuple):
    """Per-camera constants; the 12 keyword fields of gaussian_renderer/__init__.py:930-943."""
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


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def preprocess_gaussians(self, means3D, scales, rotations, shs, opacities, cuda_args=None):
        out = _ops.preprocess_gaussians(means3D, scales, rotations, shs, opacities, self.raster_settings, cuda_args)
        if self.raster_settings.debug:
            torch.cuda.synchronize()
        return out

    def render_gaussians(self, means2D, conic_opacity, rgb, depths, radii, compute_locally,
                         extended_compute_locally=None, cuda_args=None):
        out = _ops.render_gaussians(means2D, conic_opacity, rgb, depths, radii, compute_locally, self.raster_settings,
                                    cuda_args, extended_compute_locally)
        if self.raster_settings.debug:
            torch.cuda.synchronize()
        return out

    def forward(self, means3D, scales, rotations, shs, opacities, compute_locally=None, cuda_args=None):
        """Convenience: preprocess then render on one rank."""
        means2D, rgb, conic_opacity, radii, depths = self.preprocess_gaussians(means3D, scales, rotations, shs,
                                                                               opacities, cuda_args)
        image, *_ = self.render_gaussians(means2D, conic_opacity, rgb, depths, radii, compute_locally, None, cuda_args)
        return image, radii


# legacy tile-exchange operators (module level in the reference's extension; loss_distribution.py:168-195)
load_image_tiles_by_pos = _ops.load_image_tiles_by_pos
merge_image_tiles_by_pos = _ops.merge_image_tiles_by_pos
