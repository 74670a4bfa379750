"""The hand-off between the two hot paths: the Gaussians that VicaSplat.forward returns go STRAIGHT into
DecoderSplattingCUDA.forward, as in the reference's training / test steps (model_wrapper.py:207-220).
What is tested is the layout contract -- `means` as a strided view of raw_gaussians, covariances (...,3,3)
next to the packed cov6, harmonics (...,3,25) consumed without a transpose, opacities with a trailing
singleton -- by rendering the SAME predicted tensors with the CPU oracle rasterizer after re-deriving them
from raw_gaussians with the oracle's adapter (oracle.encoder_ref.gaussian_adapter -> oracle.raster_ref).
The weights are 'trained-like': random-init weights put every Gaussian within 0.16 of the camera (culled),
so the centre head's output bias is moved to z ~ 3 and the scale bias raised until the splats are a few
pixels wide."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder_ref as er
from oracle import raster_ref as rr
from oracle.make_encoder_golden import CASES, synth_inputs


def test_encoder_gaussians_render_through_the_decoder_plugin(cuda, lib):
    from vicasplat_b200 import decoder as dec
    from vicasplat_b200.encoder import VicaSplat, VicaSplatCfg, default_backbone_cfg
    kw, _, T, _ = CASES["small"]
    cfg = er.EncoderConfig(**kw)
    bb = dict(default_backbone_cfg(), img_size=cfg.img_size, enc_depth=cfg.enc_depth, dec_depth=cfg.dec_depth)
    sd = er.synth_state_dict(cfg, seed=0)
    # trained-like output statistics (see the module docstring)
    sd["downstream_head1.dpt.head.4.bias"] = torch.tensor([0.0, 0.0, 1.4])          # expm1(1.4) ~ 3 in front
    b = sd["gaussian_param_head.dpt.head.4.bias"].clone()
    b[1:4] = 4.0                                                                     # scales ~ 0.004
    sd["gaussian_param_head.dpt.head.4.bias"] = b
    model = VicaSplat(VicaSplatCfg(backbone=bb)).to(cuda).eval()
    model.load_state_dict(sd, strict=True)
    B, S, V = 2, cfg.img_size, 3
    image, K = synth_inputs(B, T, S)
    with torch.no_grad():
        out = model({"image": image.to(cuda), "intrinsics": K.to(cuda)}, compute_viewspace_depth=False)
    g = out["gaussians"]
    assert g.means.data_ptr() == out["raw_gaussians"].data_ptr() and not g.means.is_contiguous()   # a view, as upstream
    ext = torch.eye(4).repeat(B, V, 1, 1)
    ext[:, :, 0, 3] = torch.tensor([-0.2, 0.0, 0.25])
    ext[1, :, 1, 3] = 0.1                                   # the two scenes are seen from different cameras
    Kt = K[:, :1].expand(B, V, 3, 3).contiguous()
    near, far = torch.full((B, V), 0.01), torch.full((B, V), 100.0)
    decoder = dec.DecoderSplattingCUDA(dec.DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], False)).to(cuda)
    with torch.no_grad():
        ren = decoder.forward(g, ext.to(cuda), Kt.to(cuda), near.to(cuda), far.to(cuda), (S, S))
    assert ren.color.shape == (B, V, 3, S, S) and ren.depth.shape == (B, V, S, S)
    # the oracle's adapter on the SAME raw_gaussians, then the oracle rasterizer, scene by scene
    raw = out["raw_gaussians"].cpu()
    ga = er.gaussian_adapter(raw, cfg)
    assert torch.allclose(g.covariances.cpu(), ga["covariances"], rtol=1e-4, atol=1e-9)
    assert torch.allclose(g.harmonics.cpu(), ga["harmonics"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(g.opacities.cpu(), ga["opacities"], rtol=1e-5, atol=1e-6)
    worst = 0.0
    for b_ in range(B):
        flat = lambda t: t[b_].flatten(0, 2)
        rc, rd = rr.render_cuda_ref(ext[b_], Kt[b_], near[b_], far[b_], (S, S), torch.zeros((V, 3)),
                                    flat(ga["means"]), flat(ga["covariances"]), flat(ga["harmonics"]),
                                    flat(ga["opacities"])[..., 0])
        ec = (ren.color[b_].cpu() - rc).abs()
        ed = (ren.depth[b_].cpu() - rd).abs() / rd.abs().clamp_min(1)
        assert rc.abs().max() > 0.05, "the scene must be visible"          # not an empty render
        assert (ec <= 1e-4).float().mean() >= 0.995 and (ed <= 1e-4).float().mean() >= 0.995, \
            (ec.max().item(), (ec <= 1e-4).float().mean().item())
        worst = max(worst, ec.max().item())
    print(f"[hand-off] encoder Gaussians -> decoder plugin vs oracle adapter + oracle rasterizer: max colour "
          f"error {worst:.3e}")
    # scene isolation: scene 1's render does not change when scene 0's Gaussians do
    g2 = type(g)(means=g.means.clone(), covariances=g.covariances, harmonics=g.harmonics, opacities=g.opacities,
                 cov6=g.cov6)
    g2.means[0] += 10.0
    with torch.no_grad():
        ren2 = decoder.forward(g2, ext.to(cuda), Kt.to(cuda), near.to(cuda), far.to(cuda), (S, S))
    assert torch.equal(ren2.color[1], ren.color[1]) and not torch.equal(ren2.color[0], ren.color[0])
