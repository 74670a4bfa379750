"""LossMse (src/loss/loss_mse.py:23-31) through vs_mse_loss: value and gradient against plain torch
fp32/fp64, determinism, and the chain into the rasterizer's backward."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 3, 3, 17, 19), (1, 12, 3, 256, 256), (7,), (4, 4)])
def test_mse_value_and_grad(cuda, lib, shape):
    from vicasplat_b200.loss import LossMse, LossMseCfg, LossMseCfgWrapper
    g = torch.Generator().manual_seed(len(shape))
    pred = torch.rand(shape, generator=g).to(cuda).requires_grad_(True)
    tgt = torch.rand(shape, generator=g).to(cuda)
    loss_fn = LossMse(LossMseCfgWrapper(LossMseCfg(weight=0.7)))
    loss = loss_fn(SimpleNamespace(color=pred), {"target": {"image": tgt}}, None, 0)
    loss.backward()
    ref_p = pred.detach().double().requires_grad_(True)
    ref = 0.7 * ((ref_p - tgt.double()) ** 2).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item()))
    assert torch.allclose(pred.grad.double(), ref_p.grad, rtol=1e-5, atol=1e-9)
    # bit-identical on repetition (fixed-order reduction)
    l2 = loss_fn(SimpleNamespace(color=pred.detach()), {"target": {"image": tgt}}, None, 0)
    assert l2.item() == loss.item()


def test_mse_into_raster_backward(cuda, lib):
    """loss(render) -> dL/dcolor -> rasterizer backward: same Gaussian gradients as torch's own MSE."""
    from vicasplat_b200 import synthetic, decoder as dec
    from vicasplat_b200.loss import mse
    from vicasplat_b200.rasterizer import rasterize_views
    S, V = 64, 2
    sc = {k: v.to(cuda) for k, v in synthetic.gaussian_scene(2, S, S, V, seed=9, n_gauss=3000).items()}
    tanfov, view_t, full_t, campos = dec._cameras(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
    cov6 = dec._cov6(sc["covariances"]).contiguous()
    tgt = torch.rand((V, 3, S, S), device=cuda)
    grads = []
    for fused in (True, False):
        means = sc["means"].clone().requires_grad_(True)
        sh = sc["harmonics"].clone().requires_grad_(True)
        color = rasterize_views(means, cov6, sc["opacities"], shs=sh, sh_degree=4, sh_layout="chan_major",
                                viewmatrix=view_t, projmatrix=full_t, campos=campos, tanfov=tanfov,
                                bg=torch.zeros((V, 3), device=cuda), H=S, W=S)[0]
        loss = mse(color, tgt, 1.0) if fused else ((color - tgt) ** 2).mean()
        loss.backward()
        grads.append((loss.item(), means.grad.clone(), sh.grad.clone()))
    assert abs(grads[0][0] - grads[1][0]) < 1e-6
    for a, b in zip(grads[0][1:], grads[1][1:]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7)
