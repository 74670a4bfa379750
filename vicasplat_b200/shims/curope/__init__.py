"""Import-name shim: top-level ``curope`` as tried first by the reference
(croco/curope/curope2d.py:6-9) -> vicasplat_b200.curope."""
from vicasplat_b200.curope import rope_2d  # noqa: F401

__all__ = ["rope_2d"]
