"""Import-name shim so that ``from gsplat.rendering import rasterization`` (cuda_splatting.py:15)
succeeds; the gsplat render path itself is outside this hot path."""
