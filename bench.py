"""bench.py — G+D seg-GAN training step throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full iteration of the reference loop body (train_seg_gan.py:188-233): generator
forward, BCEDice + content + adversarial losses, generator backward + clamp + Adam, three
discriminator forwards, discriminator backward + clamp + Adam.  Workload at N = 1: BASELINE.json
configs[1] (batch 16 x 3 x 512 x 512, bf16); at N > 1: configs[2] (batch 8 per GPU, SyncBN +
gradient all-reduce over NCCL), weak scaling.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "G+D train imgs/s @512^2 (full seg-GAN step: U-Net G + D, BCE+Dice+content+adversarial, clamp+Adam)"
UNIT = "img/s"
G_FWD_GFLOP_512 = 417.25       # per 512^2 image
D_FWD_GFLOP_512 = 49.4
STEP_GFLOP_512 = 3 * G_FWD_GFLOP_512 + 8 * D_FWD_GFLOP_512     # 1646.9 required per image per step


def synthetic_batch(batch, cin, h, w, num_classes=3, seed=1234):
    """SURVEY.md §8(d) synthetic inputs for the GPU arm: input = randn, target = (rand > 0.5), drawn in that order from one
    seeded CPU generator (the oracle's generator draws the same stream; the product arm does not import the oracle)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cin, h, w, generator=g)
    t = (torch.rand(batch, num_classes, h, w, generator=g) > 0.5).float()
    return x, t


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (debug)")
    ap.add_argument("--size", type=int, default=512, help="tile size override (debug; headline is 512)")
    ap.add_argument("--conv", default="auto", choices=["auto", "simt"])
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--infer-batch", type=int, default=64, help="BASELINE configs[4] leg: eval-mode segmentation batch (0 = skip)")
    ap.add_argument("--cpu-sample", default="4x512", help="cpu_baseline sample BxS (bounded)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step(batch, size, steps=1, warmup=0):
    """The reference's CPU implementation of the step (oracle port of train_seg_gan.py:188-233), all host threads."""
    import functools
    import torch
    import ssunet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd_g = O.portable_state_dict(O.unet_r_ss_v2_spec(3, 3, prefix="net."))
    sd_d = O.portable_state_dict(O.discriminator_spec(3))
    og = O.AdamState(O.trainable_keys(sd_g), 2e-5)
    od = O.AdamState(O.trainable_keys(sd_d), 2e-5)
    orig = O.unet_r_ss_v2
    O.unet_r_ss_v2 = functools.partial(orig, prefix="net.")
    times = []
    try:
        for it in range(warmup + steps):
            x, t = O.synthetic_batch(batch, 3, size, size, seed=1234 + it)
            t0 = time.perf_counter()
            O.gan_train_step(sd_g, sd_d, og, od, x, t)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    finally:
        O.unet_r_ss_v2 = orig
    return times


def run_reference(args):
    """--impl reference: the reference path on the host CPU.  The reference is pure Python/PyTorch and is not
    installable on the GPU box (no /root/reference there), so the oracle port is timed (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    b, s = [int(v) for v in args.cpu_sample.split("x")]
    steps = max(1, min(args.steps, 2))
    warm = 1 if args.warmup > 0 else 0
    times = cpu_reference_step(b, s, steps=steps, warmup=warm)
    sec = sum(times) / len(times)
    # scale the bounded sample to the metric's unit: images of 512^2 per second (work is proportional to pixels)
    scale = (s * s) / (512.0 * 512.0)
    val = b * scale / sec
    cores = os.cpu_count() or 1
    sample = "full G+D step, batch %d x 3 x %d x %d fp32, %d timed step(s), pixel-scaled to 512^2" % (b, s, s, steps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "seg-GAN G+D step, 16 x 3 x 512 x 512 per GPU (configs[1]); CPU arm runs a bounded sample", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    import ssunet_gan_b200 as ssg
    from ssunet_gan_b200 import _lib, batchnorm, models_seg_gan, optim, replicate, train_step

    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl(args.conv)
    size = args.size
    batch = args.batch or (16 if world == 1 else 8)
    torch.manual_seed(41)          # train_seg_gan.py:35-36; G then D
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": 3, "deep_supervision": False}).cuda().train()
    d = models_seg_gan.Discriminator(3).cuda().train()
    if world > 1:
        g = replicate.DataParallelWithCallback(batchnorm.convert_model(g))
        d = replicate.DataParallelWithCallback(batchnorm.convert_model(d))
    og = optim.FusedClampAdam(g.parameters(), lr=2e-5)
    od = optim.FusedClampAdam(d.parameters(), lr=2e-5)

    # synthetic inputs: pinned host copies for the e2e leg, device copies for the kernel leg
    n_sets = 2
    host = []
    for i in range(n_sets):
        x, t = synthetic_batch(batch, 3, size, size, seed=1234 + rank + 97 * i)
        host.append((x.pin_memory(), t.pin_memory()))
    dev = [(x.cuda(), t.cuda()) for x, t in host]

    graphed = None
    if not args.eager:
        # the whole G+D iteration (fwd, losses, bwd, NCCL all-reduces, clamp+Adam) captured once and replayed
        graphed = train_step.GraphedGanStep(g, d, og, od, (batch, 3, size, size))

    def step_dev(i):
        x, t = dev[i % n_sets]
        if graphed is not None:
            return graphed(x, t)
        return train_step.gan_train_step(g, d, og, od, x, t, with_metrics=False)

    # end-to-end leg: every step copies ITS inputs from pinned host memory and reads the loss back.  The copy of step
    # i+1 is issued on a side stream while step i computes (double-buffered device staging), as a data loader would.
    copy_stream = torch.cuda.Stream()
    staging = [(torch.empty_like(dev[0][0]), torch.empty_like(dev[0][1])) for _ in range(2)]
    staged_ev = [None, None]

    def prefetch(i):
        hx, ht = host[i % n_sets]
        sx, st_ = staging[i % 2]
        with torch.cuda.stream(copy_stream):
            sx.copy_(hx, non_blocking=True)
            st_.copy_(ht, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged_ev[i % 2] = ev

    def step_e2e(i, last):
        if staged_ev[i % 2] is None:
            prefetch(i)
        torch.cuda.current_stream().wait_event(staged_ev[i % 2])
        sx, st_ = staging[i % 2]
        staged_ev[i % 2] = None
        if graphed is not None:
            graphed.load(sx, st_)
            if not last:
                prefetch(i + 1)            # H2D of the next batch overlaps this step's kernels
            r = graphed.replay()
        else:
            x, t = sx.clone(), st_.clone()
            if not last:
                prefetch(i + 1)
            r = train_step.gan_train_step(g, d, og, od, x, t, with_metrics=False)
        return float(r["loss"])        # D2H read of the step's result (syncs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms

    for i in range(args.warmup):
        step_dev(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms = timed(step_dev, args.steps)
    launches = graphed.launches_per_step * args.steps if graphed is not None else (_lib.launch_count - l0)
    ms_e2e = timed(lambda i: step_e2e(i, i == args.steps - 1), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # roofline leg: the dominant kernels (tcgen05 convolutions) timed one by one with CUDA events on the launching stream,
    # live, over eager steps of the same workload (a captured graph cannot carry per-kernel events)
    prof_steps = 2
    PROF = ["ssg_conv2d_fwd_tc", "ssg_conv2d_dgrad_tc", "ssg_conv2d_dgrad_tc_acc", "ssg_conv2d_dgrad_tc_split", "ssg_conv2d_wgrad_tc", "ssg_conv2d_wgrad_tc_acc"]
    train_step.gan_train_step(g, d, og, od, dev[0][0], dev[0][1], with_metrics=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.profile_reset(PROF)
    e0.record()
    for i in range(prof_steps):
        x, t = dev[i % n_sets]
        train_step.gan_train_step(g, d, og, od, x, t, with_metrics=False)
    e1.record()
    prof = _lib.profile_collect()
    ms_prof = e0.elapsed_time(e1)

    # SyncBN stat-reduce latency (BASELINE.json metric): the [sum x | sum x^2] fp64 all-reduce of one BN layer over NVLink,
    # timed alone on the compute stream (CUDA events, 50 back-to-back reduces after 10 warm-ups), smallest and largest layer
    stat_reduce_us = None
    if world > 1:
        from ssunet_gan_b200 import ops as _ops
        peer = _ops.PeerStatReducer.for_group(dist.group.WORLD)
        stat_reduce_us = {"path_used_by_the_step": "nvlink peer-memory kernel (csrc/p2p.cu)" if peer is not None else "nccl all-reduce"}

        def _time_reduce(fn, buf):
            for _ in range(10):
                fn(buf)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(50):
                fn(buf)
            a1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([a0.elapsed_time(a1) / 50 * 1e3], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return round(float(tt), 2)

        for c in (64, 768):
            buf = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
            key = "C=%d (%d B)" % (c, 16 * c)
            stat_reduce_us[key] = {"nccl": _time_reduce(lambda b: dist.all_reduce(b), buf)}
            if peer is not None:
                stat_reduce_us[key]["peer_memory_kernel"] = _time_reduce(peer.all_reduce, buf)

    # BASELINE configs[4] (extra key, not the headline): inference-only segmentation, eval-mode BN, logits -> IoU / Dice
    # through metrics.py's kernels (host round trip for the two scalars included), batch 64 x 3 x size^2 resident in HBM
    infer = None
    if world == 1 and args.infer_batch > 0:
        from ssunet_gan_b200 import metrics
        g.eval()
        ib = args.infer_batch
        xi = torch.randn(ib, 3, size, size, device="cuda")
        ti = (torch.rand(ib, 3, size, size, device="cuda") > 0.5).float()

        def infer_step(_i):
            with torch.no_grad():
                lo = g(xi)
                return metrics.iou_score(lo[:, 1:].contiguous(), ti[:, 1:].contiguous())

        for i in range(2):
            infer_step(i)
        ms_inf = timed(infer_step, 3)
        infer = {"value": ib * 3 / (ms_inf / 1e3), "unit": "img/s", "batch": ib, "ms_per_batch": ms_inf / 3,
                 "what": "Generator.eval() forward + iou_score, batch %d x 3 x %d x %d bf16" % (ib, size, size)}
        g.train()

    # Weak-scaling anchor (extra key, not the headline): BASELINE configs[1] runs batch 16 on one GPU but configs[2] runs batch 8 PER
    # GPU on 2/4/8 GPUs, so value(N) / (N * value(1)) compares different per-GPU work.  The same step at batch 8 on this one GPU is
    # the per-GPU work of the N > 1 runs: value(N) / (N * anchor) is the scaling efficiency at equal work per GPU.
    anchor = None
    if world == 1 and not args.batch and graphed is not None and batch > 8:
        b8 = 8
        xs = [(x[:b8].contiguous(), t[:b8].contiguous()) for x, t in dev]
        g8 = train_step.GraphedGanStep(g, d, og, od, (b8, 3, size, size))
        for i in range(2):
            g8(*xs[i % n_sets])
        ms8 = timed(lambda i: g8(*xs[i % n_sets]), args.steps)
        anchor = {"value": b8 * args.steps * (size * size) / (512.0 * 512.0) / (ms8 / 1e3), "unit": UNIT, "batch_per_gpu": b8,
                  "ms_per_step": ms8 / args.steps,
                  "what": "the same captured G+D step at batch 8 on one GPU = the per-GPU work of BASELINE configs[2] (N = 2/4/8)"}
        del g8

    imgs = world * batch * args.steps
    scale = (size * size) / (512.0 * 512.0)
    value = imgs * scale / (ms / 1e3)
    e2e = imgs * scale / (ms_e2e / 1e3)

    if rank != 0:
        _finish(world)
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    roof = {"bound": "tensor", "kernel": "conv2d_tc (tcgen05 implicit GEMM, fwd+dgrad+wgrad launches)", "achieved": None,
            "peak": peak_tf, "unit": "TFLOP/s", "frac": None, "traffic": None, "peak_source": peak_src}
    if prof and prof.get("ms", 0) > 0:
        ach = prof["flops"] / (prof["ms"] / 1e3) / 1e12
        roof.update({"achieved": ach, "frac": ach / peak_tf, "launches": prof["n"], "kernel_ms_per_step": prof["ms"] / prof_steps,
                     "share_of_step": prof["ms"] / ms_prof,
                     "how": "CUDA events around every tcgen05 conv launch over %d eager steps; achieved = algorithmic conv FLOPs (2*MACs, "
                            "unpadded channels) / summed launch time" % prof_steps,
                     "by_entry_point": {k: {"tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["ms"] > 0 else None,
                                            "ms_per_step": round(v["ms"] / prof_steps, 3), "launches_per_step": v["n"] // prof_steps}
                                        for k, v in sorted(prof.get("by_name", {}).items())},
                     # one `ncu --set full` capture of the most frequent instance (profiles/r01_ncu_halo_conv_l0_fwd.txt):
                     # conv0_0.conv2 forward, 16 x 512^2 x 64 -> 64: DRAM bytes vs the algorithmic read-x-once + write-y-once
                     "traffic_example": {"kernel": "conv_tc_halo_kernel<2,64,2,RES> (conv0_0.conv2 fwd)", "dram_bytes": 1.0277e9,
                                         "algorithmic_bytes": 1.0737e9}})
    step_tf = STEP_GFLOP_512 * 1e9 * scale * world * batch * args.steps / (ms / 1e3) / 1e12
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (randn tiles, Bernoulli(0.5) masks; seed-41 default init)",
            "config": {"workload": "seg-GAN G+D step: UNet_R_SS_v2 (config_v1) + SRGAN discriminator, batch %d x 3 x %d x %d per GPU%s"
                                   % (batch, size, size, "" if world == 1 else ", SyncBN + gradient all-reduce (NCCL)"),
                       "global_batch": world * batch, "parallelism": "dp%d" % world, "conv_impl": args.conv,
                       "launch": "eager (one launch per kernel)" if graphed is None else "CUDA graph replay of the captured step",
                       "l2_policy": "inputs+activations per step (GBs) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(2 * batch * 3 * size * size * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "step_tflops_required_work": step_tf, "roofline": roof, "clocks": clocks}
    if stat_reduce_us is not None:
        line["syncbn_stat_reduce_us"] = stat_reduce_us
    if infer is not None:
        line["inference"] = infer
    if anchor is not None:
        line["weak_scaling_anchor"] = anchor
    if not args.no_cpu_baseline and world == 1:
        b, s = [int(v) for v in args.cpu_sample.split("x")]
        t = cpu_reference_step(b, s, steps=1, warmup=0)
        v = b * (s * s) / (512.0 * 512.0) / t[0]
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "one full G+D step, batch %d x 3 x %d x %d fp32 (%.1f s), pixel-scaled to 512^2" % (b, s, s, t[0])}
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Leave without tearing NCCL down: destroying a communicator whose collectives live in a still-referenced CUDA graph
    can block forever (observed: the N = 2 run printed its line and then hung in destroy_process_group)."""
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


if __name__ == "__main__":
    main()
