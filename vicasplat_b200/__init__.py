"""vicasplat_b200 -- B200-native (sm_100a) hot path of VicaSplat behind the reference's plugin API.

Sub-modules:
  _lib         ctypes binding of the C-ABI (include/vicasplat_b200.h)
  ops          thin tensor-level wrappers of the C-ABI entry points
  curope       drop-in for the reference's ``curope`` extension (rope_2d, cuRoPE2D)
  rasterizer   drop-in for ``diff_gaussian_rasterization`` (GaussianRasterizationSettings/-Rasterizer)
  decoder      ``render_cuda`` / ``DecoderSplattingCUDA`` (src/model/decoder/*)
  encoder      ``VicaSplat`` encoder with the reference's state_dict keys (src/model/encoder/vicasplat.py)
  pipeline     ``ScenePipeline``: host clip in, host renders / poses out, copies overlapped with compute
  loss, optim, pose_align   ``LossMse``, ``FusedAdamW`` (torch.optim.AdamW state layout), test-time pose alignment
  encoder_grad hand-written forward-for-training / backward of the ViT block (croco/blocks.py:81-130)
  blocks       drop-in trainable ``Block`` on those kernels (gradients through torch.autograd)
  encoder_train ``VitEncoderTrainer``: image-encoder training step, ``GradBucket`` / ``GradReducer``
               (bucketed, overlapped gradient all-reduce), ``plan_buckets`` for the whole model
  synthetic    seeded synthetic clips, Gaussian scenes and weights of BASELINE.json's shapes
"""
__version__ = "0.1.0"
