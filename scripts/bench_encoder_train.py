"""Training step of the ViT-L image encoder at the size of BASELINE configs[2]/[3]:
`scenes` 8-view 256x256 clips per GPU and step; forward + hand-written backward + bucketed gradient
all-reduce (N > 1) + fused AdamW + bf16 repack.  One process per GPU:

    python scripts/bench_encoder_train.py --scenes 8 --steps 5 --warmup 2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 scripts/bench_encoder_train.py --scenes 8

Rank 0 prints one JSON line (device-timed, max over ranks).  The gradient of the encoder output is
synthetic (the decoder / heads backward is not built yet): this measures the encoder's share of the
step and the overlap of its all-reduce, not a full VicaSplat training step.
"""
import argparse
import json
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from vicasplat_b200 import dist_util, synthetic
from vicasplat_b200.encoder_train import GradReducer, ViTEncoderConfig, VitEncoderTrainer
from vicasplat_b200.optim import FusedAdamW


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--depth", type=int, default=24)
    ap.add_argument("--no-reduce", action="store_true", help="skip the all-reduce (N > 1): its cost by difference")
    ap.add_argument("--bf16-reduce", action="store_true", help="all-reduce bf16 copies of the gradient buckets")
    a = ap.parse_args()
    rank, world, local = dist_util.rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = ViTEncoderConfig(enc_depth=a.depth)
    sd = synthetic.vit_encoder_state_dict(depth=a.depth, seed=0)
    frames = a.scenes * 8
    reducer = GradReducer(compress_bf16=a.bf16_reduce)
    if a.no_reduce:
        reducer.world = 1
    tr = VitEncoderTrainer(sd, cfg, frames, (256, 256), dev, reducer=reducer)
    del sd
    opt = FusedAdamW(tr.parameters(), lr=2e-5, betas=(0.9, 0.95), weight_decay=0.05, max_grad_norm=0.5)
    g = torch.Generator().manual_seed(dist_util.scene_seed(250307, rank))
    img = (torch.rand((frames, 3, 256, 256), generator=g) * 2 - 1).to(dev)
    K = torch.tensor([[0.86, 0, 0.5], [0, 0.86, 0.5], [0, 0, 1.0]]).repeat(frames, 1, 1).to(dev)
    d_out = (torch.randn((frames * tr.lay.n, cfg.enc_embed_dim), generator=g) * 1e-3).to(torch.bfloat16).to(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    parts = {"fwd": 0.0, "bwd": 0.0, "opt": 0.0}

    def step(record):
        e = [ev() for _ in range(4)]
        e[0].record()
        tr.forward(img, K)
        e[1].record()
        tr.backward(d_out)
        e[2].record()
        opt.step()
        tr.repack()
        e[3].record()
        if record is not None:
            record.append(e)

    for _ in range(a.warmup):
        step(None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    rec = []
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(a.steps):
        step(rec)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = t0.elapsed_time(t1) / a.steps
    for e in rec:
        parts["fwd"] += e[0].elapsed_time(e[1]) / a.steps
        parts["bwd"] += e[1].elapsed_time(e[2]) / a.steps
        parts["opt"] += e[2].elapsed_time(e[3]) / a.steps
    red = dist_util.max_over_ranks([ms, parts["fwd"], parts["bwd"], parts["opt"]], dev)
    if rank == 0:
        n_params = sum(p.numel() for p in tr.parameters())
        M, E = frames * tr.lay.n, cfg.enc_embed_dim
        lin = 2.0 * M * 12 * E * E * a.depth + 2.0 * frames * 256 * 768 * E
        att = 4.0 * tr.lay.n ** 2 * 64 * cfg.enc_num_heads * frames * a.depth
        flop = 3 * lin + 3.5 * att                       # fwd + dgrad + wgrad; attention fwd + 2.5x bwd
        print(json.dumps({
            "metric": "ViT-L encoder training step (E1-E3 of SURVEY §8), scenes/s", "unit": "scenes/s",
            "value": round(a.scenes * world / (red[0] / 1e3), 2), "n_gpus": world, "scaling": "weak",
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(red[0], 3),
            "fwd_ms": round(red[1], 3), "bwd_incl_allreduce_ms": round(red[2], 3),
            "adamw_repack_ms": round(red[3], 3), "scenes_per_step_per_gpu": a.scenes,
            "params": n_params, "allreduce_bytes_per_step": reducer.bytes_reduced // max(1, a.steps + a.warmup),
            "algorithmic_tflops": round(flop / red[0] / 1e9, 1), "dtype": "bf16 operands, fp32 accumulate / master",
            "data": "synthetic (random-init weights, seeded images, synthetic output gradient)",
            "reduce": not a.no_reduce, "reduce_dtype": "bf16" if a.bf16_reduce else "fp32",
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
