"""FusedAdamW (vs_adamw_step) against torch.optim.AdamW + clip_grad_norm_ with the reference's settings
(betas (0.9, 0.95), weight_decay 0.05, two learning-rate groups, clip 0.5: model_wrapper.py:884-951,
config/main.yaml:70)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(cuda, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1024, 768), (768,), (3, 3, 64, 17), (5,), (40000,), (1,)]
    return [torch.randn(s, generator=g).to(cuda).requires_grad_(True) for s in shapes]


def test_matches_torch_adamw_with_clipping(cuda, lib):
    from vicasplat_b200.optim import FusedAdamW
    ours, ref = _make(cuda, 0), _make(cuda, 0)
    kw = dict(weight_decay=0.05, betas=(0.9, 0.95), eps=1e-8)
    o = FusedAdamW([{"params": ours[:3], "lr": 4e-5}, {"params": ours[3:], "lr": 4e-6}], lr=4e-5,
                   max_grad_norm=0.5, **kw)
    r = torch.optim.AdamW([{"params": ref[:3], "lr": 4e-5}, {"params": ref[3:], "lr": 4e-6}], lr=4e-5, **kw)
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        scale = 10.0 if step % 2 == 0 else 1e-3        # clipped and un-clipped steps
        for a, b in zip(ours, ref):
            gr = (torch.randn(a.shape, generator=g) * scale).to(cuda)
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 3:
            o.param_groups[0]["lr"] = r.param_groups[0]["lr"] = 2e-5    # scheduler changed the lr
        want_norm = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        r.step()
        o.step()
        assert abs(o.grad_norm.item() - want_norm.item()) <= 1e-5 * want_norm.item()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (step, (a - b).abs().max())
    for a, b in zip(ours, ref):
        assert torch.allclose(o.state[a]["exp_avg_sq"], r.state[b]["exp_avg_sq"], rtol=1e-5, atol=1e-12)
        assert torch.allclose(o.state[a]["exp_avg"], r.state[b]["exp_avg"], rtol=1e-5, atol=1e-9)


def test_nonfinite_gradient_skips_the_step(cuda, lib):
    from vicasplat_b200.optim import FusedAdamW
    ps = _make(cuda, 2)
    o = FusedAdamW(ps, lr=1e-3, max_grad_norm=0.5, nonfinite="skip")
    before = [p.detach().clone() for p in ps]
    for p in ps:
        p.grad = torch.ones_like(p)
    ps[2].grad[0, 0, 0, 0] = float("nan")
    o.step()
    assert o.found_inf.item() == 1
    assert all(torch.equal(a, b) for a, b in zip(ps, before))
    assert o.step_count == 0                      # a skipped step does not advance the bias corrections
    ps[2].grad[0, 0, 0, 0] = 0.0
    o.step()
    assert o.found_inf.item() == 0 and not torch.equal(ps[0], before[0])
    assert o.step_count == 1 and float(o.state_dict()["state"][0]["step"]) == 1.0


def test_nonfinite_gradients_are_sanitised_like_the_reference(cuda, lib):
    """default: GradientNanCheckCallback (src/main.py:40-45) = torch.nan_to_num_ on the gradients, then the
    usual clip + AdamW step."""
    from vicasplat_b200.optim import FusedAdamW
    ours, ref = _make(cuda, 3), _make(cuda, 3)
    kw = dict(weight_decay=0.05, betas=(0.9, 0.95), eps=1e-8)
    o = FusedAdamW(ours, lr=1e-3, max_grad_norm=0.5, **kw)
    r = torch.optim.AdamW(ref, lr=1e-3, **kw)
    g = torch.Generator().manual_seed(4)
    for step in range(3):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).to(cuda)
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 1:
            for t in (ours[0].grad, ref[0].grad):
                t[3, 5] = float("nan")
                t[7, 1] = float("nan")
        for b in ref:                              # the reference's callback
            if torch.isnan(b.grad).any():
                torch.nan_to_num_(b.grad)
        torch.nn.utils.clip_grad_norm_(ref, 0.5)
        r.step()
        o.step()
        assert o.found_inf.item() == (1 if step == 1 else 0)
        for a, b in zip(ours, ref):
            assert torch.isfinite(a).all() and torch.allclose(a, b, rtol=2e-6, atol=1e-7), step
    assert o.step_count == 3
