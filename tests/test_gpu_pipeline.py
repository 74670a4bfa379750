"""ScenePipeline (copy / compute overlap over consecutive batches) returns exactly what the
sequential plugin calls return."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_matches_sequential_calls(cuda, lib):
    from vicasplat_b200 import synthetic
    from vicasplat_b200 import decoder as dec
    from vicasplat_b200.encoder import Gaussians, VicaSplat, VicaSplatCfg, default_backbone_cfg
    from vicasplat_b200.pipeline import ScenePipeline
    torch.manual_seed(3)
    bb = dict(default_backbone_cfg(), enc_depth=2, dec_depth=10, img_size=64)
    model = VicaSplat(VicaSplatCfg(backbone=bb)).to(cuda).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "modulation" in n or n.startswith("camera_extrinsic_head"):
                p.normal_(0, 0.02)
    model.invalidate()
    decoder = dec.DecoderSplattingCUDA(dec.DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], False)).to(cuda)
    B, T, S, V = 2, 4, 64, 3
    sc = {k: v.to(cuda) for k, v in synthetic.gaussian_scene(T, S, S, V, seed=5).items()}
    rep = lambda t: t[None].expand(B, *t.shape)
    gauss = Gaussians(means=rep(sc["means"]), covariances=rep(sc["covariances"]),
                      harmonics=rep(sc["harmonics"]), opacities=rep(sc["opacities"]))
    target = dict(extrinsics=rep(sc["extrinsics"]), intrinsics=rep(sc["intrinsics"]), near=rep(sc["near"]),
                  far=rep(sc["far"]), image_shape=(S, S), gaussians=gauss)
    clips = [synthetic.clip(B, T, S, seed=100 + i) for i in range(5)]
    clips = [(im.pin_memory(), K.pin_memory()) for im, K in clips]

    ref = []
    for im, K in clips:
        enc = model({"image": im.to(cuda), "intrinsics": K.to(cuda)}, compute_viewspace_depth=False)
        o = decoder.forward(gauss, target["extrinsics"], target["intrinsics"], target["near"], target["far"], (S, S))
        ref.append((o.color.cpu(), o.depth.cpu(), enc["pred_extrins"].cpu()))

    pipe = ScenePipeline(model, decoder, depth=2)
    got, pending = [], []
    for im, K in clips:
        pending.append(pipe.submit({"image": im, "intrinsics": K}, target))
        if len(pending) == 2:                       # collect batch i-1 while batch i is in flight
            r = pending.pop(0).result()
            got.append({k: v.clone() for k, v in r.items()})
    for t in pending:
        r = t.result()
        got.append({k: v.clone() for k, v in r.items()})
    pipe.drain()
    assert len(got) == len(ref)
    for (c, d, p), r in zip(ref, got):
        assert torch.equal(c, r["color"]) and torch.equal(d, r["depth"]) and torch.equal(p, r["pred_extrins"])
    # poses differ between clips (the pipeline really ran each batch)
    assert not torch.equal(got[0]["pred_extrins"], got[1]["pred_extrins"])


def test_stale_ticket_raises(cuda, lib):
    from vicasplat_b200.pipeline import Ticket, _Slot
    s = _Slot()
    s.seq = 5
    with pytest.raises(RuntimeError):
        Ticket(s, 3).result()
