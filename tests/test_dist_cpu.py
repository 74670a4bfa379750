"""world_size-2 gloo test of the N>1 plumbing (replicas + max-over-ranks timing) and of the
reference arm's rank gating."""
import json
import os
import subprocess
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    from vicasplat_b200 import dist_util
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, lr = dist_util.rank_world()
    ms = [10.0 + 5 * rank, 3.0 - rank]                      # rank 1 slower on [0], faster on [1]
    red = dist_util.max_over_ranks(ms, "cpu")
    seeds = [None, None]
    dist.all_gather_object(seeds, dist_util.scene_seed(100, r))
    if rank == 0:
        out.put((r, w, red, seeds, dist_util.aggregate_throughput(4, w, red[0] / 1e3)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_max_reduce_and_seeds():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    r, w, red, seeds, thr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (r, w) == (0, 2)
    assert red == [15.0, 3.0]
    assert len(set(seeds)) == 2
    assert abs(thr - 2 * 4 / 0.015) < 1e-6


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_models_match_survey_figures():
    sys.path.insert(0, str(ROOT))
    import bench
    assert abs(bench.encoder_flops(8) / 1e9 - 3405) < 10          # 3 405 GF/scene at T=8
    assert abs(bench.encoder_flops(2) / 1e9 - 817) < 5
    assert abs(bench.raster_bytes(1, 524288, 256, 256) / 1e6 - 122.7) < 0.1   # 122.7 MB / view
    cfg = bench.config_dict(4)
    assert "workload" in cfg and cfg["scenes_per_step_per_gpu"] == 4 and "model" not in cfg
    json.dumps(cfg)


# ------------------------------------------------------------------ gradient buckets + reducer (C4's collective)
def test_grad_bucket_layout():
    sys.path.insert(0, str(ROOT))
    from vicasplat_b200.encoder_train import GradBucket
    shapes = {"b": torch.Size([6]), "g": torch.Size([10]), "w": torch.Size([6, 10])}
    bk = GradBucket(shapes, ["b", "g"], ["w"], "cpu")
    assert bk.flat.numel() == 8 + 12 + 60 and bk.n_small == 20          # 16-byte aligned views
    assert bk.views["w"].shape == (6, 10) and bk.views["w"].data_ptr() == bk.flat[20:].data_ptr()
    bk.flat.fill_(1.0)
    bk.zero(weights_too=False)
    assert bk.views["b"].abs().sum() == 0 and bk.views["g"].abs().sum() == 0 and bk.views["w"].sum() == 60
    bk.zero()
    assert bk.flat.abs().sum() == 0


def _reduce_worker(rank, world, port, out, compress=False):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    from vicasplat_b200.encoder_train import GradBucket, GradReducer
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = {"b": torch.Size([4]), "w": torch.Size([3, 4])}
    buckets = [GradBucket(shapes, ["b"], ["w"], "cpu") for _ in range(3)]
    red = GradReducer(compress_bf16=compress)
    for i in reversed(range(3)):                              # reverse layer order, like the backward pass
        buckets[i].views["b"].fill_(float(rank + 1) * (i + 1))
        buckets[i].views["w"].copy_(torch.arange(12.0).view(3, 4) * (rank + 1))
        red.bucket_ready(buckets[i])
    red.finish()
    if rank == 0:
        out.put(([b.views["b"].tolist() for b in buckets], buckets[1].views["w"].tolist(),
                 red.world, red.bytes_reduced))
    dist.barrier()
    dist.destroy_process_group()


@__import__("pytest").mark.parametrize("compress", [False, True])
def test_two_rank_gloo_gradient_buckets_are_averaged(compress):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + (7 if compress else 0)
    procs = [ctx.Process(target=_reduce_worker, args=(r, 2, port, q, compress)) for r in range(2)]
    for p in procs:
        p.start()
    bs, w, world, nbytes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert world == 2 and nbytes == 3 * 16 * (2 if compress else 4)      # bf16 on the wire halves the bytes
    assert bs == [[1.5 * (i + 1)] * 4 for i in range(3)]                 # mean of rank 1x and 2x
    assert w == (torch.arange(12.0).view(3, 4) * 1.5).tolist()


def test_reducer_is_a_noop_without_a_process_group():
    sys.path.insert(0, str(ROOT))
    from vicasplat_b200.encoder_train import GradBucket, GradReducer
    bk = GradBucket({"b": torch.Size([4])}, ["b"], [], "cpu")
    bk.flat.fill_(2.0)
    red = GradReducer()
    red.bucket_ready(bk)
    red.finish()
    assert red.world == 1 and bk.flat.tolist() == [2.0] * 4


def test_bucket_plan_covers_every_parameter_and_skips_the_dead_ones():
    """C4's static plan against the reference: the parameters left out of the buckets are exactly the
    ones the unmodified reference never gives a gradient (tests/golden/unused_params.json, measured
    by oracle/make_unused_params_golden.py), everything else sits in exactly one bucket, and buckets
    come in reverse execution order."""
    sys.path.insert(0, str(ROOT))
    from oracle import encoder_ref as er
    from vicasplat_b200.encoder import VicaSplat, VicaSplatCfg, default_backbone_cfg
    from vicasplat_b200.encoder_train import _BLOCK_BIG, _BLOCK_SMALL, plan_buckets
    gold = json.loads((ROOT / "tests" / "golden" / "unused_params.json").read_text())
    bb = dict(default_backbone_cfg(), img_size=64, enc_depth=2, dec_depth=10)
    model = VicaSplat(VicaSplatCfg(backbone=bb))
    names = [k for k, _ in model.named_parameters()]
    assert len(names) == gold["n_parameters"]
    buckets, unused = plan_buckets(names, enc_depth=2, dec_depth=10)
    assert sorted(unused) == gold["no_grad"] and gold["zero_grad"] == []
    flat = [k for _, ks in buckets for k in ks]
    assert len(flat) == len(set(flat)) and set(flat) | set(unused) == set(names)
    order = [n for n, _ in buckets]
    assert order[:4] == ["gs_head", "pts_head", "cam_head", "dec_norms"]
    assert order[4:15] == [f"dec{i}" for i in range(9, -1, -1)] + ["dec_stem"]
    assert order[15:] == ["enc1", "enc0", "enc_stem"]
    enc1 = dict(buckets)["enc1"]
    assert sorted(enc1) == sorted(f"backbone.enc_blocks.1.{n}" for n in _BLOCK_SMALL + _BLOCK_BIG)
    with __import__("pytest").raises(KeyError):
        plan_buckets(["backbone.enc_blocks.7.norm1.weight"], enc_depth=2, dec_depth=10)
