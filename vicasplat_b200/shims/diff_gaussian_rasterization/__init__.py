"""Import-name shim: ``diff_gaussian_rasterization`` as the reference imports it
(src/model/decoder/cuda_splatting.py:5-8) -> vicasplat_b200.rasterizer."""
from vicasplat_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
