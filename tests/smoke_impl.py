"""One small invocation of the hot path on cuda:0, checked against the oracle."""
import torch


def run_smoke():
    from oracle import raster_ref as rr
    from vicasplat_b200.decoder import render_cuda
    dev = torch.device("cuda:0")
    hw = 48
    sc = rr.synthetic_scene(1, hw, hw, 2, seed=4)
    rc, rd = rr.render_cuda_ref(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"],
                                (hw, hw), torch.zeros((2, 3)), sc["means"],
                                sc["covariances"], sc["harmonics"], sc["opacities"])
    d = {k: v.to(dev) for k, v in sc.items()}
    c, dep = render_cuda(d["extrinsics"], d["intrinsics"], d["near"], d["far"], (hw, hw),
                         torch.zeros((2, 3), device=dev), d["means"], d["covariances"],
                         d["harmonics"], d["opacities"])
    torch.cuda.synchronize()
    ec = (c.cpu().double() - rc).abs()
    ok = (ec <= 1e-4).double().mean().item()
    assert ok > 0.999, f"raster smoke mismatch: {ok}"
    print(f"smoke: raster ok ({ok:.5f} of pixels within 1e-4, max {ec.max():.2e})")
