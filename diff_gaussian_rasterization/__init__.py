"""Drop-in import name.  `from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer` (gs-simp/gaussian_renderer/__init__.py:14 of the reference) resolves to the
B200-native implementation in multiview_inpaint_b200."""
from multiview_inpaint_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                                _RasterizeGaussians, rasterize_gaussians)
from multiview_inpaint_b200 import _C  # noqa: F401
