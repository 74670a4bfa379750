"""Checkpoint / resume contract of FusedAdamW (host logic, no GPU): its state round-trips with
torch.optim.AdamW's state_dict layout, the optimizer the reference checkpoints
(src/model/model_wrapper.py:884-951)."""
import pytest
import torch


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).requires_grad_(True) for s in [(6, 4), (4,), (3, 2, 2)]]


def _torch_opt(ps):
    return torch.optim.AdamW([{"params": ps[:2], "lr": 4e-5}, {"params": ps[2:], "lr": 4e-6}],
                             lr=4e-5, betas=(0.9, 0.95), weight_decay=0.05)


def test_state_round_trips_with_torch_adamw():
    from vicasplat_b200.optim import FusedAdamW
    ref = _params(0)
    topt = _torch_opt(ref)
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        for p in ref:
            p.grad = torch.randn(p.shape, generator=g)
        topt.step()
    ours = _params(0)
    fopt = FusedAdamW([{"params": ours[:2], "lr": 1.0}, {"params": ours[2:], "lr": 1.0}], lr=1.0,
                      betas=(0.9, 0.95), weight_decay=0.05)
    ptr_before = [fopt.state[p]["exp_avg"].data_ptr() for p in ours]
    fopt.load_state_dict(topt.state_dict())
    assert fopt.step_count == 3
    assert [g["lr"] for g in fopt.param_groups] == [4e-5, 4e-6]
    assert ptr_before == [fopt.state[p]["exp_avg"].data_ptr() for p in ours]    # copied in place
    for a, b in zip(ours, ref):
        assert torch.equal(fopt.state[a]["exp_avg"], topt.state[b]["exp_avg"])
        assert torch.equal(fopt.state[a]["exp_avg_sq"], topt.state[b]["exp_avg_sq"])
    # and back: torch's AdamW accepts what FusedAdamW saves
    again = _torch_opt(_params(0))
    again.load_state_dict(fopt.state_dict())
    for (_, a), (_, b) in zip(sorted(again.state_dict()["state"].items()), sorted(topt.state_dict()["state"].items())):
        assert float(a["step"]) == float(b["step"]) == 3.0
        assert torch.equal(a["exp_avg"], b["exp_avg"]) and torch.equal(a["exp_avg_sq"], b["exp_avg_sq"])
    assert [g["lr"] for g in again.param_groups] == [4e-5, 4e-6]


def test_fresh_state_and_mismatches():
    from vicasplat_b200.optim import FusedAdamW
    ours = _params(2)
    fopt = FusedAdamW(ours, lr=1e-3)
    fopt.load_state_dict(torch.optim.AdamW(_params(2), lr=1e-3).state_dict())     # never stepped
    assert fopt.step_count == 0
    with pytest.raises(ValueError):
        fopt.load_state_dict(_torch_opt(_params(2)).state_dict())                  # two groups vs one
    bad = fopt.state_dict()
    bad["state"][0]["step"] = torch.tensor(5.0)
    with pytest.raises(ValueError, match="disagree"):
        fopt.load_state_dict(bad)


def test_step_without_cuda_fails_loudly():
    from vicasplat_b200.optim import FusedAdamW
    ps = _params(3)
    for p in ps:
        p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdamW(ps, lr=1e-3).step()


def test_group_hyperparameters_are_validated_not_ignored():
    """one weight decay / betas / eps reaches the kernel: a group that asks for something else must raise
    (a silently decayed no-decay group would be a wrong optimizer)."""
    from vicasplat_b200.optim import FusedAdamW
    ps = _params(4)
    with pytest.raises(ValueError, match="weight_decay"):
        FusedAdamW([{"params": ps[:2]}, {"params": ps[2:], "weight_decay": 0.0}], lr=1e-3, weight_decay=0.05)
    FusedAdamW([{"params": ps[:2]}, {"params": ps[2:], "weight_decay": 0.05}], lr=1e-3, weight_decay=0.05)
    with pytest.raises(ValueError, match="no trainable"):
        FusedAdamW([], lr=1e-3)
    fopt = FusedAdamW(ps, lr=1e-3, weight_decay=0.05)
    with pytest.raises(ValueError, match="weight_decay"):
        fopt.load_state_dict(torch.optim.AdamW(_params(4), lr=1e-3, weight_decay=0.01).state_dict())
