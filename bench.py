#!/usr/bin/env python
"""Headline benchmark: scenes/s (and rendered Mpix/s) of the 8-view 256x256 forward path
(encoder + 12-view splatting render), BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); scenes are independent so ranks
run replicas of the step with no data-path collective ("weak" scaling); rank 0 prints ONE JSON line.
A step = one scene: VicaSplat encoder forward on an 8-frame clip, then the 12-view render of a
524 288-Gaussian scene.  Data is synthetic (vicasplat_b200/synthetic.py): random-weight encoder
output is degenerate as raster input, so the raster leg renders a seeded
pixel-aligned Gaussian scene of exactly the shape the encoder emits.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

T_CTX, V_TGT, SIZE = 8, 12, 256
G_SCENE = T_CTX * SIZE * SIZE
METRIC = "scenes_per_sec_8view_256x256_forward"


# ------------------------------------------------------------------------------------ models
def encoder_flops(T: int, size: int = SIZE) -> float:
    """Algorithmic forward FLOPs (2*MAC) per scene of T frames."""
    E, D, Le, Ld = 1024, 768, 24, 12
    N = (size // 16) ** 2 + 1
    px = (size / 256.0) ** 2
    per_frame = (2 * N * 12 * E * E * Le + 4 * N * N * E * Le                     # ViT-L linear + attention
                 + 2 * (N * 16 + 21) * D * D * Ld                                 # MixDecoder linears
                 + 4 * N * T * (N + 1) * D * Ld                                   # video attention
                 + 4 * N * (2 * N if T > 2 else N) * D * Ld                       # neighbour attention
                 + 2 * (N - 1) * 768 * E + 2 * N * E * D                          # patch / decoder embed
                 + (62.2e9 + 118.2e9) * px)                                       # two DPT heads
    return per_frame * T


def outconv_flops_saved(T: int, size: int = SIZE) -> float:
    """FLOPs NOT executed per scene: the four 1x1 out_convs of both DPT trunks run before their
    bilinear x2 (they commute), i.e. on a quarter of the pixels the reference convolves."""
    g = size // 16
    px = sum((g * s) ** 2 for s in (1, 2, 4, 8))          # out_conv output pixels per frame, per head
    return T * 2 * px * 2 * 256 * 256 * 0.75


def attention_flops(T: int, size: int = SIZE) -> float:
    """The part of encoder_flops done by the attention kernel (QK^T and PV), per scene."""
    E, D, Le, Ld = 1024, 768, 24, 12
    N = (size // 16) ** 2 + 1
    return T * (4 * N * N * E * Le + 4 * N * T * (N + 1) * D * Ld + 4 * N * (2 * N if T > 2 else N) * D * Ld)


# DRAM traffic of all GEMM launches of one scene's encoder forward, from the committed ncu capture
# (profiles/r1_kernel_traffic.txt: sum of dram__bytes_read + dram__bytes_write); None = not captured
GEMM_DRAM_BYTES_PER_SCENE = 4.79e9      # 1-scene capture; 1.16e9 of it are the bf16 weights
GEMM_WEIGHT_BYTES = 1.16e9               # read once per launch whatever the batch


def raster_bytes(V: int, G: int, H: int, W: int) -> float:
    """Algorithmic HBM bytes of a V-view forward render: every used Gaussian
    parameter once per view (12 B mean + 24 B cov6 + 4 B opacity + 192 B of SH bands 0..3) plus one
    RGB-D write per pixel."""
    return V * (G * 232.0 + H * W * 16.0)


def load_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], which="measured")
    return dict(hbm=6650.0, tf=1400.0, which="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.result = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            self.result = dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx),
                               reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------ reference arm
def cpu_reference_step(t_frames: int, n_views: int, seed: int = 1):
    """One bounded sample of the workload on the host cores with the oracle port of the reference
    (the reference itself cannot run here: /root/reference does not exist on the GPU box and its
    rasterizer is an un-vendored CUDA-only extension).  Returns (t_encoder, t_raster_per_view)."""
    import torch
    from oracle import encoder_ref as er
    from oracle import raster_ref as rr
    from vicasplat_b200 import synthetic
    # all host cores (torchrun exports OMP_NUM_THREADS=1 for its workers)
    if torch.get_num_threads() < (os.cpu_count() or 1):
        torch.set_num_threads(os.cpu_count() or 1)
    st = cpu_reference_step
    if not hasattr(st, "cache"):
        cfg = er.EncoderConfig()
        st.cache = dict(cfg=cfg, sd=er.synth_state_dict(cfg, seed=0),
                        scene=synthetic.gaussian_scene(T_CTX, SIZE, SIZE, V_TGT, seed=seed))
    c = st.cache
    image, K = synthetic.clip(1, t_frames, SIZE)
    with torch.no_grad():
        t0 = time.perf_counter()
        er.forward(c["sd"], image, K, c["cfg"])
        t1 = time.perf_counter()
        sc = c["scene"]
        rr.render_cuda_ref(sc["extrinsics"][:n_views], sc["intrinsics"][:n_views], sc["near"][:n_views],
                           sc["far"][:n_views], (SIZE, SIZE), torch.zeros((n_views, 3)), sc["means"],
                           sc["covariances"], sc["harmonics"], sc["opacities"])
        t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) / n_views


def torch_eager_gpu_encoder(dev, our_ms_per_scene: float) -> dict:
    import torch
    from oracle import encoder_ref as er
    from vicasplat_b200 import synthetic
    c = cpu_reference_step.cache
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    try:
        sd = {k: v.to(dev) for k, v in c["sd"].items()}
        image, K = synthetic.clip(1, T_CTX, SIZE)
        image, K = image.to(dev), K.to(dev)
        with torch.no_grad():
            for _ in range(2):
                er.forward(sd, image, K, c["cfg"])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                er.forward(sd, image, K, c["cfg"])
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return dict(encoder_ms_per_scene=ms, ours_encoder_ms_per_scene=our_ms_per_scene,
                encoder_speedup=ms / our_ms_per_scene,
                what="oracle port of the encoder (functional torch, fp32 weights, TF32 matmul/conv, eager, "
                     "1 scene of 8 frames) on this GPU; encoder leg only")


def run_reference(args) -> None:
    """The reference's CPU path (the oracle port: the reference itself needs 16 absent packages and a
    CUDA-only rasterizer) on all host cores.  Every step is ONE WHOLE scene, timed as it runs -- the full
    8-frame encoder forward and all 12 target views -- nothing is extrapolated: `ms_per_step` is the
    measured wall time of a step and `value` = 1 scene / that.  (Our arm runs `--batch` such scenes per
    step; scenes are independent, so scenes/s is the same quantity.)  One scene takes ~20 s on 8 cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sample = (f"per step: one whole scene = oracle encoder forward on all {T_CTX} frames + all {V_TGT} target "
              f"views of the {G_SCENE}-Gaussian scene, wall-clock, no extrapolation (1 of the {args.batch} "
              f"scenes of a step of our arm)")
    for _ in range(args.warmup):
        cpu_reference_step(T_CTX, V_TGT)
    secs = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_reference_step(T_CTX, V_TGT)
        secs.append(time.perf_counter() - t0)
    per = sum(secs) / len(secs)
    val = 1.0 / per
    line = dict(impl="reference", metric=METRIC, value=val, unit="scenes/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=per * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", extrapolated=False,
                scenes_per_step=1,
                mpix_per_sec=val * V_TGT * SIZE * SIZE / 1e6,
                config=config_dict(args.batch),
                raster_parity="unpinned (upstream diff_gaussian_rasterization source absent; restated algorithm)",
                cpu_baseline=dict(value=val, unit="scenes/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=val, unit="scenes/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def config_dict(nb: int = 1) -> dict:
    return dict(workload=f"configs[1]: {T_CTX}-view {SIZE}x{SIZE} re10k_8view forward-only, {nb} scene(s) per "
                         f"GPU per step: encoder ({T_CTX} frames/scene) + render of {V_TGT} target views/scene, "
                         f"G={G_SCENE} Gaussians/scene", scenes_per_step_per_gpu=nb, context_views=T_CTX,
                target_views=V_TGT,
                image=f"{SIZE}x{SIZE}", gaussians=G_SCENE, weights="random-init (seeded)",
                raster_input="one DIFFERENT seeded pixel-aligned Gaussian scene per scene of the step "
                             "(random-weight encoder output is degenerate as raster input)",
                l2="inputs larger than L2 (1.2 GB bf16 weights, 178 MB Gaussians); no explicit flush",
                parallelism="replicas (one scene per GPU, no data-path collective)")


# ------------------------------------------------------------------------------------ our arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    from vicasplat_b200 import dist_util
    rank, world, local = dist_util.rank_world()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from vicasplat_b200 import _lib, synthetic
    from vicasplat_b200 import decoder as dec
    from vicasplat_b200.encoder import Gaussians, VicaSplat
    from vicasplat_b200.rasterizer import rasterize_views
    lib = _lib.load()

    torch.manual_seed(1234 + rank)
    model = VicaSplat().to(dev).eval()
    with torch.no_grad():                      # the reference zero-inits these; exercise them
        for n, p in model.named_parameters():
            if "modulation" in n or n.startswith("camera_extrinsic_head"):
                p.normal_(0, 0.02)
    model.invalidate()
    eng = model.engine()
    NB = args.batch
    image_h, K_h = synthetic.clip(NB, T_CTX, SIZE, seed=dist_util.scene_seed(250307, rank))
    image_h, K_h = image_h.pin_memory(), K_h.pin_memory()
    image_d, K_d = image_h.to(dev), K_h.to(dev)
    # NB DIFFERENT scenes per step (scene 0 of rank 0 is the CPU-generated one the reference arm renders);
    # cov6 is what the adapter writes next to the covariances (encoder.Gaussians.cov6)
    scenes = []
    for i in range(NB):
        seed = dist_util.scene_seed(1 + 104729 * i, rank)
        sc_i = synthetic.gaussian_scene(T_CTX, SIZE, SIZE, V_TGT, seed=seed, device=None if i == 0 else dev)
        sc_i = {k: v.to(dev) for k, v in sc_i.items()}
        sc_i["cov6"] = dec._cov6(sc_i["covariances"]).contiguous()
        scenes.append(sc_i)
    sc = scenes[0]
    bg = torch.zeros((V_TGT, 3), device=dev)
    ext_all = torch.stack([s_["extrinsics"] for s_ in scenes]).flatten(0, 1)
    intr_all = torch.stack([s_["intrinsics"] for s_ in scenes]).flatten(0, 1)
    near_all = torch.stack([s_["near"] for s_ in scenes]).flatten()
    far_all = torch.stack([s_["far"] for s_ in scenes]).flatten()
    rkw = dict(sh_degree=4, sh_layout="chan_major", bg=bg, H=SIZE, W=SIZE,
               want_n_touched=False)   # render_cuda discards it (cuda_splatting.py:226-239)
    import vicasplat_b200.rasterizer as rmod

    def raster_batch(check="deferred"):
        # the camera set-up of all NB * V cameras (sync-free) belongs to the step
        tanfov, view_t, full_t, campos = dec._cameras(ext_all, intr_all, near_all, far_all)
        with rmod.SceneStreams(dev) as ss:        # one launch chain (all 12 views) per scene, scenes
            for i, s_ in enumerate(scenes):       # round-robin over SceneStreams (what the decoder plugin does)
                c = slice(i * V_TGT, (i + 1) * V_TGT)
                with ss.scene(i):
                    out = rasterize_views(s_["means"], s_["cov6"], s_["opacities"], shs=s_["harmonics"],
                                          viewmatrix=view_t[c], projmatrix=full_t[c], campos=campos[c],
                                          tanfov=tanfov[c], check_overflow=check, **rkw)
                ss.keep(*out)
        return out

    def verify_raster():
        """the deferred overflow counters of every render since the last call (one small D2H copy)"""
        recs = rmod.take_deferred()
        if recs:
            rmod.verify_deferred(torch.stack([r[0] for r in recs]).cpu(), recs)

    # calibrate the binning capacities once, outside the timed region: checked calls read the exact
    # pair counts / largest per-tile counts from device scalars and leave them (+25 %) in the
    # rasterizer's per-shape hint; the timed calls use the hint and DEFER the check of their own
    # counters (the shipped path of ScenePipeline), verified right after the timed region
    for _ in range(2):
        color = raster_batch(check=True)[0]
    max_pairs, max_tile = rmod._capacity_hint[(V_TGT, G_SCENE, SIZE, SIZE)]
    n_pairs = int((max_pairs - 4096) / 1.25)

    def device_step():
        eng.run(image_d, K_d, clone_outputs=False)
        return raster_batch()

    # launches per step: kernels recorded into the encoder graph + the raster chain
    l0 = lib.vs_launch_count()
    device_step()                               # first call captures the graph (counts twice: warm-up + capture)
    torch.cuda.synchronize()
    l1 = lib.vs_launch_count()
    raster_batch()
    l2 = lib.vs_launch_count()
    raster_launches = l2 - l1
    enc_launches = (l1 - l0 - raster_launches) // 2
    gpu_launches = int(enc_launches + raster_launches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    verify_raster()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            ev[i][0].record()
            eng.run(image_d, K_d, clone_outputs=False)
            ev[i][1].record()
            raster_batch()
            ev[i][2].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    verify_raster()
    if os.environ.get("VS_PROFILE_STEP") == "1":
        # one extra step between cudaProfilerStart/Stop (outside the timed region) for
        #   ncu --profile-from-start off --metrics gpu__time_duration.sum ... python bench.py
        torch.cuda.cudart().cudaProfilerStart()
        device_step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    seq_ms = ev[0][0].elapsed_time(ev[-1][2]) / args.steps
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    ras_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    step_ms, clocks = seq_ms, clk.result

    if args.pipelined:
        # Steady-state pipeline: the render of scene batch i (stream B, waits for the encoder of
        # batch i) runs while the encoder of batch i+1 occupies the tensor cores (stream A).
        sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
        def pipelined(n):
            for _ in range(n):
                with torch.cuda.stream(sA):
                    eng.run(image_d, K_d, clone_outputs=False)
                    done = torch.cuda.Event()
                    done.record()
                with torch.cuda.stream(sB):
                    sB.wait_event(done)
                    raster_batch()
        cur = torch.cuda.current_stream()
        sA.wait_stream(cur); sB.wait_stream(cur)
        pipelined(3)
        cur.wait_stream(sA); cur.wait_stream(sB)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk2:
            e0.record()
            sA.wait_stream(cur); sB.wait_stream(cur)
            pipelined(args.steps)
            cur.wait_stream(sA); cur.wait_stream(sB)
            e1.record()
            barrier()
        step_ms, clocks = e0.elapsed_time(e1) / args.steps, clk2.result

    # ---- roofline leg: one un-graphed encoder pass with CUDA events around every launch of the
    # dominant kernel family (the graph replays exactly these launches)
    from vicasplat_b200 import ops as vops
    from vicasplat_b200.encoder import EncoderEngine
    eager = EncoderEngine(model, use_graph=False)
    eager.run(image_d, K_d, clone_outputs=False)
    vops.TIMERS = {}
    eager.run(image_d, K_d, clone_outputs=False)
    fam = vops.family_ms(vops.TIMERS)
    n_gemm = len(vops.TIMERS.get("gemm", []))
    vops.TIMERS = None
    del eager
    gemm_ms = fam.get("gemm", float("nan"))

    # ---- e2e: host buffers in, host result out, through the plugin calls a user makes
    decoder = dec.DecoderSplattingCUDA(dec.DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], False)).to(dev)
    verify_raster()
    stk = lambda k: torch.stack([s_[k] for s_ in scenes])
    gauss = Gaussians(means=stk("means"), covariances=stk("covariances"), harmonics=stk("harmonics"),
                      opacities=stk("opacities"), cov6=stk("cov6"))
    ext, intr = stk("extrinsics"), stk("intrinsics")
    near, far = stk("near"), stk("far")
    # Every step submits one host batch (pinned clip -> H2D) and collects the host result of the
    # previous one (colour, depth, poses <- D2H): all copies of all K steps happen inside the
    # timed region; the pipeline only overlaps them with the compute of the neighbouring steps.
    from vicasplat_b200.pipeline import ScenePipeline
    pipe = ScenePipeline(model, decoder, depth=2)
    target = dict(extrinsics=ext, intrinsics=intr, near=near, far=far, image_shape=(SIZE, SIZE),
                  gaussians=gauss)
    host_ctx = {"image": image_h, "intrinsics": K_h}
    sink = torch.zeros((), dtype=torch.float64)

    def e2e_steps(n):
        prev = None
        for _ in range(n):
            t = pipe.submit(host_ctx, target)
            if prev is not None:
                r = prev.result()
                sink.add_(float(r["pred_extrins"][0, 0, 0]) + float(r["color"][0, 0, 0, 0, 0]))
            prev = t
        r = prev.result()                      # the last batch is drained inside the timed region too
        sink.add_(float(r["pred_extrins"][0, 0, 0]) + float(r["color"][0, 0, 0, 0, 0]))

    e2e_steps(3)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    slot = pipe.slots[0]
    h2d = image_h.numel() * 4 + K_h.numel() * 4
    d2h = (slot.color_h.numel() + slot.depth_h.numel() + slot.pose_h.numel()) * 4

    # ---- max over ranks
    step_ms, enc_ms, ras_ms, e2e_ms, seq_ms = dist_util.max_over_ranks(
        [step_ms, enc_ms, ras_ms, e2e_s * 1e3, seq_ms], dev)

    # overflow check of the calibrated capacities (outside the timed region): a checked call would
    # have re-run and changed the hint if either bound had been exceeded
    chk = raster_batch(check=True)[0]
    assert torch.isfinite(chk).all() and torch.equal(chk, color)

    # ---- parity mode (precision="fp16": TF32's mantissa, see VicaSplat): the same encoder step, timed alike
    from vicasplat_b200.encoder import EncoderEngine as _Eng
    eng16 = _Eng(model, precision="fp16")
    for _ in range(3):
        eng16.run(image_d, K_d, clone_outputs=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng16.run(image_d, K_d, clone_outputs=False)
    e1.record()
    barrier()
    enc16_ms, = dist_util.max_over_ranks([e0.elapsed_time(e1) / args.steps], dev)
    del eng16

    stress = None
    if args.stress_steps > 0:
        stress = stress_leg(args, dev, rank, world, local, barrier)
    train = None
    if args.train_batch > 0:
        train = train_leg(args, model, scenes, dev, rank, world, local, barrier)

    if rank == 0:
        peaks = load_peaks()
        value = world * NB * 1e3 / step_ms
        fl = NB * encoder_flops(T_CTX)
        enc_tf = fl / (enc_ms * 1e-3) / 1e12
        gemm_fl = NB * (encoder_flops(T_CTX) - attention_flops(T_CTX) - outconv_flops_saved(T_CTX))   # executed
        gemm_tf = gemm_fl / (gemm_ms * 1e-3) / 1e12
        rb = NB * raster_bytes(V_TGT, G_SCENE, SIZE, SIZE)
        ras_gbs = rb / (ras_ms * 1e-3) / 1e9
        line = dict(
            metric=METRIC, value=value, unit="scenes/s", n_gpus=world, steps=args.steps,
            warmup=max(args.warmup, 3), ms_per_step=step_ms, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="bf16", data="synthetic", config=config_dict(NB),
            mpix_per_sec=value * V_TGT * SIZE * SIZE / 1e6,
            encoder_ms=enc_ms, raster_ms=ras_ms, sequential_ms_per_step=seq_ms, raster_pairs=n_pairs,
            raster_sort=("per-tile shared-memory sort" if max_tile > 0 else "global 64-bit radix sort"),
            schedule=("sequential" if not args.pipelined else
                      "pipelined: render of batch i (stream B, event-dependent on encoder i) overlaps "
                      "encoder of batch i+1 (stream A); encoder_ms / raster_ms are un-overlapped"),
            e2e=dict(value=world * NB * 1e3 / e2e_ms, unit="scenes/s", h2d_bytes_per_step=h2d,
                     d2h_bytes_per_step=d2h, ms_per_step=e2e_ms,
                     api="vicasplat_b200.pipeline.ScenePipeline.submit/result over VicaSplat.forward + "
                         "DecoderSplattingCUDA.forward: pinned host clip in, host colour/depth/poses out, "
                         "copies of step i overlap compute of steps i+-1 (all inside the timed region)"),
            gpu_launches=gpu_launches,
            clocks=clocks,
            roofline=dict(bound="tensor", kernel="gemm_tc05_kernel",
                          achieved=gemm_tf, peak=peaks["tf"], unit="TFLOP/s", frac=gemm_tf / peaks["tf"],
                          traffic=(GEMM_WEIGHT_BYTES + NB * (GEMM_DRAM_BYTES_PER_SCENE - GEMM_WEIGHT_BYTES))
                          / max(n_gemm, 1),
                          peak_source=peaks["which"] + " bf16 sustained",
                          launches_per_step=n_gemm, flops_per_launch=gemm_fl / max(n_gemm, 1),
                          ms_per_step=gemm_ms, share_of_encoder=gemm_ms / enc_ms,
                          note="sum over the GEMM / implicit-GEMM launches of one encoder forward, CUDA events "
                               "around each launch in an un-graphed pass; traffic = ncu dram bytes per launch "
                               "(profiles/, scaled by scenes per step)"),
            roofline_encoder=dict(bound="tensor", kernel="whole encoder forward (all kernels, graph replay)",
                                  achieved=enc_tf, peak=peaks["tf"], unit="TFLOP/s",
                                  frac=enc_tf / peaks["tf"], flops_per_step=fl),
            roofline_raster=dict(bound="hbm", kernel="preprocess+sort+blend chain, 12 views",
                                 achieved=ras_gbs, peak=peaks["hbm"], unit="GB/s",
                                 frac=ras_gbs / peaks["hbm"], traffic=None,
                                 peak_source=peaks["which"] + " copy", bytes_per_launch=rb),
            wall_ms_per_step=t_wall * 1e3 / args.steps,
            raster_parity="unpinned (upstream diff_gaussian_rasterization source absent; restated algorithm)",
        )
        line["parity_mode"] = dict(
            precision="fp16 operands (10-bit mantissa = TF32's, the reference's matmul precision), fp32 accumulation "
                      "and residual streams", encoder_ms=enc16_ms,
            value=world * NB * 1e3 / (enc16_ms + ras_ms), unit="scenes/s",
            parity="raw Gaussians 1.1e-3 rel-L2 vs the reference's fp32 golden at full depth "
                   "(tests/test_gpu_encoder.py; speed mode: 1.1e-2)")
        if stress is not None:
            line["raster_stress"] = stress
        if train is not None:
            line["train_step"] = train
        if world == 1 and not args.no_cpu:
            te, tv = cpu_reference_step(T_CTX, V_TGT)
            per = te + tv * V_TGT
            line["cpu_baseline"] = dict(
                value=1.0 / per, unit="scenes/s", cores=torch.get_num_threads(), kind="port",
                sample=f"one whole scene, nothing extrapolated: oracle encoder forward on all {T_CTX} frames "
                       f"({te:.1f} s) + all {V_TGT} target views ({tv * V_TGT:.1f} s)")
            # context for the encoder leg (SURVEY.md §8d "its own PyTorch path"): the same oracle
            # port in eager PyTorch on THIS GPU with TF32 matmuls/convs, as the reference runs
            # (backbone_vica.py:9) -- a baseline measurement by the checker, nothing we ship
            try:
                line["cpu_baseline"]["torch_eager_gpu"] = torch_eager_gpu_encoder(dev, enc_ms / NB)
            except Exception as e:   # never let the context measurement break the bench line
                line["cpu_baseline"]["torch_eager_gpu"] = dict(error=str(e)[:200])
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ C5 stress leg
def stress_leg(args, dev, rank, world, local, barrier) -> dict:
    """BASELINE configs[4]: 16-view 512x512 raster-only forward of a 2^21-Gaussian scene per GPU (replicas)."""
    import torch
    from vicasplat_b200 import decoder as dec, dist_util, synthetic
    from vicasplat_b200.rasterizer import rasterize_views
    import vicasplat_b200.rasterizer as rmod
    T, V, S, G = 16, 16, 512, 1 << 21
    sc = synthetic.gaussian_scene(T, S, S, V, seed=dist_util.scene_seed(4, rank), n_gauss=G, device=dev)
    tanfov, view_t, full_t, campos = dec._cameras(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
    cov6 = dec._cov6(sc["covariances"]).contiguous()
    kw = dict(shs=sc["harmonics"], sh_degree=4, sh_layout="chan_major", viewmatrix=view_t, projmatrix=full_t,
              campos=campos, tanfov=tanfov, bg=torch.zeros((V, 3), device=dev), H=S, W=S, want_n_touched=False)
    for _ in range(2):
        ref = rasterize_views(sc["means"], cov6, sc["opacities"], check_overflow=True, **kw)[0]
    n_pairs = int((rmod._capacity_hint[(V, G, S, S)][0] - 4096) / 1.25)
    K = args.stress_steps
    for _ in range(3):
        rasterize_views(sc["means"], cov6, sc["opacities"], check_overflow="deferred", **kw)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(K):
            out = rasterize_views(sc["means"], cov6, sc["opacities"], check_overflow="deferred", **kw)[0]
        e1.record()
        barrier()
    recs = rmod.take_deferred()
    rmod.verify_deferred(torch.stack([r[0] for r in recs]).cpu(), recs)
    assert torch.equal(out, ref)
    ms, = dist_util.max_over_ranks([e0.elapsed_time(e1) / K], dev)
    peaks = load_peaks()
    gbs = raster_bytes(V, G, S, S) / (ms * 1e-3) / 1e9
    return dict(metric="raster_only_scenes_per_sec_16view_512x512", value=world * 1e3 / ms, unit="scenes/s",
                ms_per_step=ms, steps=K, warmup=5, n_gpus=world, mpix_per_sec=world * 1e3 / ms * V * S * S / 1e6,
                workload="configs[4]: 16-view 512x512 raster-only forward, G = 2^21 Gaussians per scene, one scene "
                         "per GPU per step", raster_pairs=n_pairs, clocks=clk.result,
                roofline=dict(bound="hbm", kernel="preprocess+sort+blend chain, 16 views", achieved=gbs,
                              peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                              note="SURVEY 8d per-view bytes; the chain is issue-bound, not HBM-bound (profiles/)"))


# ------------------------------------------------------------------------------------ training leg
def train_leg(args, model, scenes, dev, rank, world, local, barrier) -> dict:
    """BASELINE configs[2] (1 GPU) / configs[3] (N GPUs): one optimisation step of the 8-view 256x256
    training loop per GPU -- encoder forward (activations kept), 12-view render + MSE value/gradient +
    rasterizer backward per scene, hand-written encoder backward with the gradient all-reduce of ALL
    578 M parameters overlapped bucket by bucket (NCCL; the path's one collective, src/main.py:110-115),
    nan_to_num / clip 0.5 / fused AdamW, operand re-pack.  Batch = --train-batch scenes per GPU in
    micro-batches of --train-micro (gradient accumulation is exact).
    Raster input: the step's synthetic pixel-aligned scenes PLUS the encoder's own outputs as residuals
    (means, cov6, SH; opacity offset by -0.5), unit Jacobian: random-init weights put every predicted
    Gaussian within 0.16 of the camera where the rasterizer culls it, which would time an empty render."""
    import torch
    import torch.distributed as dist
    from vicasplat_b200 import _lib, dist_util
    from vicasplat_b200.encoder_train import GradReducer
    from vicasplat_b200.rasterizer import RasterOverflow
    from vicasplat_b200.train_step import TrainStep, _NoReduce
    lib = _lib.load()
    B, mb, K = args.train_batch, min(args.train_micro, args.train_batch), args.train_steps
    NS = len(scenes)
    model.train()
    reducer = GradReducer(compress_bf16=args.train_wire == "bf16")
    lp = None
    if args.train_lpips == "stand-in":
        from vicasplat_b200.lpips import LpipsVgg
        lp = LpipsVgg.stand_in(dev, seed=0)
    ts = TrainStep(model, micro_batch=mb, reducer=reducer, lpips=lp)
    image_h, K_h = __import__("vicasplat_b200.synthetic", fromlist=["clip"]).clip(
        B, T_CTX, SIZE, seed=dist_util.scene_seed(777, rank))
    g = torch.Generator().manual_seed(dist_util.scene_seed(778, rank))
    tgt_h = torch.rand((B, V_TGT, 3, SIZE, SIZE), generator=g)
    image_h, K_h, tgt_h = image_h.pin_memory(), K_h.pin_memory(), tgt_h.pin_memory()
    pick = lambda k: torch.stack([scenes[b % NS][k] for b in range(B)])
    target = dict(extrinsics=pick("extrinsics"), intrinsics=pick("intrinsics"), near=pick("near"), far=pick("far"))

    def override(b, gz):
        s_ = scenes[b % NS]
        return dict(means=s_["means"] + gz["means"], cov6=s_["cov6"] + gz["cov6"], sh=s_["harmonics"] + gz["sh"],
                    opac=s_["opacities"] + (gz["opac"] - 0.5))

    ctx_ext = torch.eye(4, device=dev).repeat(B, T_CTX, 1, 1)       # context cameras on the unit baseline
    ctx_ext[:, :, 0, 3] = torch.arange(T_CTX, device=dev) / (T_CTX - 1)

    def one(check, host=False):
        if host:
            ctx = dict(image=image_h.to(dev, non_blocking=True), intrinsics=K_h.to(dev, non_blocking=True),
                       extrinsics=ctx_ext)
            tgt = dict(target, image=tgt_h.to(dev, non_blocking=True))
        else:
            ctx, tgt = ctx_d, tgt_d
        loss = ts.step(ctx, tgt, override_gaussians=override, check_overflow=check)
        return float(loss) if host else loss

    ctx_d = dict(image=image_h.to(dev), intrinsics=K_h.to(dev), extrinsics=ctx_ext)
    tgt_d = dict(target, image=tgt_h.to(dev))
    for _ in range(3):                       # first steps: calibrate the rasterizer's capacity hints
        try:
            one(True)
        except RasterOverflow:
            pass
    l0 = lib.vs_launch_count()
    loss0 = one(True)
    launches = int(lib.vs_launch_count() - l0)

    def timed(n, host=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            loss = one(True, host)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, loss

    with ClockSampler(local) as clk:
        ms, _, loss = timed(K)
    e2e_ms = timed(K, host=True)[1]
    exposed = None
    if world > 1:                            # the same step without the collective: what the all-reduce costs
        ts.eng.reducer = ts.reducer = _NoReduce()
        ms_no = timed(K)[0]
        ts.eng.reducer = ts.reducer = reducer
        ms_no, = dist_util.max_over_ranks([ms_no], dev)
    ms, e2e_ms = dist_util.max_over_ranks([ms, e2e_ms], dev)
    if world > 1:
        exposed = ms - ms_no
    loss = float(loss)
    assert loss == loss and abs(loss) < 1e6, f"training loss is not finite: {loss}"
    peaks = load_peaks()
    n_params = sum(p.numel() for p in ts.eng.trainable_parameters())
    tf = 3.0 * encoder_flops(T_CTX) * B / (ms * 1e-3) / 1e12
    model.eval()
    return dict(
        metric="train_scenes_per_sec_8view_256x256", value=world * B * 1e3 / ms, unit="scenes/s",
        ms_per_step=ms, steps=K, warmup=4, scenes_per_step_per_gpu=B, micro_batch=mb, n_gpus=world,
        peak_memory_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30,
        workload=(f"configs[{2 if world == 1 else 3}]: 8-view 256x256 training step, batch {B}/GPU in micro-batches "
                  f"of {mb}: encoder fwd + 12-view render + MSE + raster bwd + encoder bwd (all {n_params} "
                  "trained parameters) + " + ("gradient all-reduce + " if world > 1 else "") +
                  "nan_to_num / clip 0.5 / AdamW + re-pack"),
        loss=dict(kind="MSE (weight 1) + dual-quaternion camera loss (weight 0.1) + LPIPS (weight 0.05) " +
                       ("with a seeded RANDOM VGG16 of the real architecture (the lpips package's weights are not in "
                        "the image): real cost, not a perceptual metric" if lp is not None else "switched off"),
                  first=float(loss0),
                  last=loss),
        raster_input="synthetic pixel-aligned scenes + encoder outputs as residuals (unit Jacobian)",
        clocks=clk.result, gpu_launches=launches,
        e2e=dict(value=world * B * 1e3 / e2e_ms, unit="scenes/s", ms_per_step=e2e_ms,
                 h2d_bytes_per_step=(image_h.numel() + K_h.numel() + tgt_h.numel()) * 4, d2h_bytes_per_step=4,
                 api="TrainStep.step: pinned host clips + target images in, host loss out, per step"),
        allreduce=dict(bytes_per_step=(sum(b_.flat.numel() for b_ in ts.eng.buckets.values()) *
                                       (2 if args.train_wire == "bf16" else 4)) if world > 1 else 0,
                       wire_dtype=args.train_wire if world > 1 else None, buckets=len(ts.eng.buckets),
                       exposed_ms=exposed, ms_per_step_without=(ms - exposed) if exposed is not None else None),
        roofline=dict(bound="tensor", kernel="whole training step (3 x forward FLOPs of the encoder; the render "
                                             "chain and optimizer are HBM-bound and not counted)",
                      achieved=tf, peak=peaks["tf"], unit="TFLOP/s", frac=tf / peaks["tf"]))


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main() -> None:
    # stdout carries exactly one JSON line: anything a library prints there (NCCL writes its
    # "NCCL version ..." banner to stdout when the communicator is created) goes to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="scenes per step per GPU")
    ap.add_argument("--pipelined", action="store_true",
                    help="overlap the render of batch i with the encoder of batch i+1 on two streams "
                         "(measured: +2 %, the persistent GEMM CTAs own the SMs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--train-batch", type=int, default=24,
                    help="scenes per GPU per TRAINING step of the train_step sub-record (BASELINE configs[2]/[3]: "
                         "24); 0 = skip the training leg")
    ap.add_argument("--train-micro", type=int, default=12,
                    help="scenes per micro-batch of the training step (12 keeps ~6 GB of activations per scene "
                         "within 180 GB; 24 does not fit)")
    ap.add_argument("--train-steps", type=int, default=3)
    ap.add_argument("--train-lpips", choices=["stand-in", "off"], default="stand-in",
                    help="LPIPS term of the training step: a seeded random VGG16 (the real weights are not in the image)")
    ap.add_argument("--stress-steps", type=int, default=5,
                    help="timed steps of the raster_stress sub-record (BASELINE configs[4]); 0 = skip")
    ap.add_argument("--train-wire", choices=["f32", "bf16"], default="bf16",
                    help="wire format of the gradient all-reduce (bf16: half the bytes; buckets stay fp32)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
