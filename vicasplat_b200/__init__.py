"""vicasplat_b200 -- B200-native (sm_100a) hot path of VicaSplat behind the reference's plugin API.

Sub-modules:
  _lib         ctypes binding of the C-ABI (include/vicasplat_b200.h)
  ops          thin tensor-level wrappers of the C-ABI entry points
  curope       drop-in for the reference's ``curope`` extension (rope_2d, cuRoPE2D)
  rasterizer   drop-in for ``diff_gaussian_rasterization`` (GaussianRasterizationSettings/-Rasterizer)
  decoder      ``render_cuda`` / ``DecoderSplattingCUDA`` (src/model/decoder/*)
  encoder      ``VicaSplat`` encoder with the reference's state_dict keys (src/model/encoder/vicasplat.py)
"""
__version__ = "0.1.0"
