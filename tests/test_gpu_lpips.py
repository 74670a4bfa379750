"""The LPIPS consumer (vicasplat_b200.lpips: VGG16 features on the implicit-GEMM conv kernel, max pooling, the
per-layer distance with its gradient) against the torch restatement oracle/lpips_ref.py with the same seeded
stand-in weights (the lpips package and its weights are not in the image: parity with the package itself is
UNPINNED, see the oracle's header).  bf16 operands / fp32 accumulation."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import lpips_ref as lr


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def test_maxpool_forward_backward(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = F.relu(torch.randn((2, 8, 12, 64), generator=g)).to(torch.bfloat16).to(cuda)
    y = ops.maxpool2(x)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(y.float(), yr.detach().permute(0, 2, 3, 1))
    dy = torch.randn(y.shape, generator=g).to(torch.bfloat16).to(cuda)
    add = torch.randn(x.shape, generator=g).to(torch.bfloat16).to(cuda)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    got = ops.maxpool2_backward(x, y, dy, add=add)
    # windows whose maximum is tied (zeros after the ReLU) may route to a different element than torch
    ref = xr.grad.permute(0, 2, 3, 1) + add.float()
    same = (got.float() - ref).abs() <= 1e-2 * ref.abs() + 1e-2
    assert same.float().mean() > 0.97
    masked = ops.maxpool2_backward(x, y, dy, relu_mask=True)
    assert torch.count_nonzero(masked[x == 0]) == 0


def test_lpips_layer_value_and_gradient(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(1)
    n, h, w, c = 3, 8, 8, 128
    f0 = F.relu(torch.randn((n, h, w, c), generator=g)).to(torch.bfloat16).to(cuda)
    f1 = F.relu(torch.randn((n, h, w, c), generator=g)).to(torch.bfloat16).to(cuda)
    wl = torch.rand((c,), generator=g).to(cuda)
    per = torch.zeros((n,), device=cuda)
    df0 = ops.lpips_layer(f0, f1, wl, per, 0.37 / (h * w))
    a = f0.float().requires_grad_(True)
    n0 = a / (a.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    n1 = f1.float() / (f1.float().pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    d = ((n0 - n1) ** 2 * wl).sum(-1).mean((1, 2))
    assert torch.allclose(per, d.detach(), rtol=1e-4, atol=1e-6)
    (0.37 * d.sum()).backward()
    assert _rel(df0, a.grad * (f0.float() > 0)) < 5e-3


def test_lpips_against_the_torch_restatement(cuda, lib):
    from vicasplat_b200.lpips import LpipsVgg
    w = LpipsVgg.stand_in_weights(seed=0)
    net = LpipsVgg(w, cuda, is_stand_in=True)
    g = torch.Generator().manual_seed(2)
    pred = torch.rand((2, 3, 64, 64), generator=g).to(cuda)
    target = (pred.cpu() + 0.1 * torch.randn((2, 3, 64, 64), generator=g)).clamp(0, 1).to(cuda)
    loss, grad = net.loss_and_grad(pred, target, weight=0.05)
    wc = {k: v.to(cuda) for k, v in w.items()}
    pr = pred.clone().requires_grad_(True)
    ref = 0.05 * lr.lpips(wc, pr, target).mean()
    ref.backward()
    print(f"[lpips] value {loss.item():.6f} vs {ref.item():.6f}; gradient rel-L2 {_rel(grad, pr.grad):.3e}, cosine "
          f"{F.cosine_similarity(grad.flatten(), pr.grad.flatten(), dim=0).item():.5f}")
    assert abs(loss.item() - ref.item()) <= 3e-2 * abs(ref.item())
    # thirteen ReLUs and four max-poolings between the loss and the image: ~1 % of the masks / argmaxes differ
    # between the bf16 and the fp32 forward pass, each flip switches a gradient path (tests/test_gpu_model_grad.py)
    assert F.cosine_similarity(grad.flatten(), pr.grad.flatten(), dim=0).item() > 0.97
    # autograd plumbing of the plugin class
    from types import SimpleNamespace
    from vicasplat_b200.lpips import LossLpips
    mod = LossLpips(SimpleNamespace(weight=0.05, apply_after_step=0), net)
    p5 = pred.view(1, 2, 3, 64, 64).clone().requires_grad_(True)
    out = mod(SimpleNamespace(color=p5), {"target": {"image": target.view(1, 2, 3, 64, 64)}})
    out.backward()
    assert torch.allclose(p5.grad.flatten(0, 1), grad)
