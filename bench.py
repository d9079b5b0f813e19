"""bench.py — G+D seg-GAN training step throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config headline|sn7]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full iteration of the reference loop body (train_seg_gan.py:188-233): generator forward, BCEDice + content +
adversarial losses, IoU / Dice, generator backward + clamp + Adam, three discriminator forwards, discriminator backward +
clamp + Adam.  Workloads:
  --config headline (default)   N = 1: BASELINE.json configs[1] (batch 16 x 3 x 512 x 512, bf16);
                                N > 1: configs[2] (batch 8 per GPU, SyncBN + gradient all-reduce over NVLink), weak scaling
  --config sn7                  configs[3]: SpaceNet7-shaped 4-band 1024 x 1024 tiles, batch 4 per GPU (any N; the config names N = 8)
Prints ONE JSON line on rank 0.  The CPU arm (`--impl reference`, and the `cpu_baseline` leg of our arm) runs the reference's
own modules from baseline/_ref (kind "reference"; staged by baseline/make_ref.py) or, when that copy is absent, the oracle port
(kind "port") -- the only places this file touches oracle/ or baseline/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT,):
    if p not in sys.path:
        sys.path.insert(0, p)

UNIT = "img/s"
G_FWD_GFLOP_512 = 417.25       # per 512^2 image
D_FWD_GFLOP_512 = 49.4
STEP_GFLOP_512 = 3 * G_FWD_GFLOP_512 + 8 * D_FWD_GFLOP_512     # 1646.9 required per image per step

CONFIGS = {
    # name: cin, tile, per-GPU batch at N = 1 / N > 1, extra conv GFLOP per image for the wider stem, BASELINE index
    "headline": dict(cin=3, size=512, batch1=16, batchn=8, extra_gflop=0.0, baseline="configs[1] (N = 1) / configs[2] (N > 1)"),
    "sn7": dict(cin=4, size=1024, batch1=4, batchn=4, extra_gflop=1.3, baseline="configs[3]"),
}


def metric_name(cfg):
    if cfg["cin"] == 3:
        return "G+D train imgs/s @512^2 (full seg-GAN step: U-Net G + D, BCE+Dice+content+adversarial, IoU/Dice, clamp+Adam)"
    return "G+D train imgs/s @%d^2 %d-band (full seg-GAN step: U-Net G + D, BCE+Dice+content+adversarial, IoU/Dice, clamp+Adam)" % (
        cfg["size"], cfg["cin"])


def synthetic_batch(batch, cin, h, w, num_classes=3, seed=1234):
    """SURVEY.md §8(d) synthetic inputs: input = randn, target = (rand > 0.5), drawn in that order from one seeded CPU
    generator (the oracle's generator draws the same stream; the product arm does not import the oracle)."""
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
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (debug)")
    ap.add_argument("--size", type=int, default=0, help="tile size override (debug)")
    ap.add_argument("--conv", default="auto", choices=["auto", "simt"])
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference leg (and the parity gate that uses it)")
    ap.add_argument("--infer-batch", type=int, default=64, help="BASELINE configs[4] leg: eval-mode segmentation batch (0 = skip)")
    ap.add_argument("--cpu-sample", default="", help="CPU arm sample BxS (bounded); default 4x512 (headline) / 1x1024 (sn7)")
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


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules (baseline/_ref) or, failing that, the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(batch, size, cin, steps=1, warmup=0):
    """Run the reference G+D step on the host CPU (all threads).  Returns dict(kind, times [s], first = result of the first
    iteration (losses, iou, dice, logits), init = (generator, discriminator) state_dicts the run started from)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_step
    if ref_step.available():
        times, first, init = ref_step.timed_steps(batch, size, input_channels=cin, steps=steps, warmup=warmup, make_batch=synthetic_batch)
        return {"kind": "reference", "times": times, "first": first, "init": init}
    # the verbatim copy is absent (it is git-ignored and staged by baseline/make_ref.py): time the oracle port instead
    import functools
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ssunet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd_g = O.portable_state_dict(O.unet_r_ss_v2_spec(3, cin, prefix="net."))
    sd_d = O.portable_state_dict(O.discriminator_spec(3))
    init = ({k: v.clone() for k, v in sd_g.items()}, {k: v.clone() for k, v in sd_d.items()})
    og = O.AdamState(O.trainable_keys(sd_g), 2e-5)
    od = O.AdamState(O.trainable_keys(sd_d), 2e-5)
    orig = O.unet_r_ss_v2
    O.unet_r_ss_v2 = functools.partial(orig, prefix="net.")
    times, first = [], None
    try:
        for it in range(warmup + steps):
            x, t = synthetic_batch(batch, cin, size, size, seed=1234 + it)
            t0 = time.perf_counter()
            r = O.gan_train_step(sd_g, sd_d, og, od, x, t)
            dt = time.perf_counter() - t0
            if it == 0:
                first = {k: r[k] for k in ("loss", "content", "adv_g", "adv_d", "iou", "dice", "logits")}
            if it >= warmup:
                times.append(dt)
    finally:
        O.unet_r_ss_v2 = orig
    return {"kind": "port", "times": times, "first": first, "init": init}


def cpu_sample_of(args, cfg):
    if args.cpu_sample:
        b, s = [int(v) for v in args.cpu_sample.split("x")]
        return b, s
    return (4, 512) if cfg["size"] <= 512 else (1, cfg["size"])


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU implementation of the step on the host cores, each step a bounded sample
    of the workload.  Rank 0 alone runs it; other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    b, s = cpu_sample_of(args, cfg)
    steps = max(1, min(args.steps, 3))      # bounded: ~8 s per step on 16 cores; the whole run ends within a minute
    warm = 1 if args.warmup > 0 else 0
    res = cpu_reference_steps(b, s, cfg["cin"], steps=steps, warmup=warm)
    sec = sum(res["times"]) / len(res["times"])
    size = args.size or cfg["size"]
    val = b * (s * s) / float(size * size) / sec          # images of the config's tile size per second (work ~ pixels)
    sample = "full G+D step, batch %d x %d x %d x %d fp32, %d timed step(s) after %d warm-up, pixel-scaled to %d^2" % (
        b, cfg["cin"], s, s, steps, warm, size)
    line = {"impl": "reference", "metric": metric_name(cfg), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(cfg, cfg["batch1"] if args.gpus == 1 else cfg["batchn"], size, args.gpus) +
                                   "; the CPU arm runs a bounded sample of it", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": res["kind"], "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(cfg, batch, size, world):
    return ("seg-GAN G+D step: UNet_R_SS_v2 (config_v1, input_channels=%d) + SRGAN discriminator, batch %d x %d x %d x %d per GPU%s [BASELINE %s]"
            % (cfg["cin"], batch, cfg["cin"], size, size, "" if world == 1 else ", SyncBN statistics over NVLink peer memory + NCCL gradient all-reduces beside the discriminator phase",
               cfg["baseline"]))


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def parity_gate(cpu, cin, b, s):
    """One EAGER bf16 step of this package from the same seed-41 weights on the same synthetic batch as the CPU reference
    step that was just timed: rel-L2 of the logits, relative error of the four losses, IoU / Dice differences."""
    import torch
    from ssunet_gan_b200 import models_seg_gan, optim, train_step
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": cin, "deep_supervision": False})
    d = models_seg_gan.Discriminator(3)
    g.load_state_dict(cpu["init"][0])
    d.load_state_dict(cpu["init"][1])
    g.cuda().train(); d.cuda().train()
    og = optim.FusedClampAdam(g.parameters(), lr=2e-5)
    od = optim.FusedClampAdam(d.parameters(), lr=2e-5)
    x, t = synthetic_batch(b, cin, s, s, seed=1234)
    r = train_step.gan_train_step(g, d, og, od, x.cuda(), t.cuda(), with_metrics=True)
    torch.cuda.synchronize()
    ref = cpu["first"]
    out = {"against": "CPU %s step, batch %d x %d x %d x %d, same seed-41 weights and inputs" % (cpu["kind"], b, cin, s, s),
           "logits_rel_l2": rel_l2(r["logits"], ref["logits"])}
    for k in ("loss", "content", "adv_g", "adv_d"):
        out[k + "_rel_err"] = abs(float(r[k]) - ref[k]) / abs(ref[k])
    out["iou_abs_err"] = abs(float(r["iou"]) - float(ref["iou"]))
    out["dice_abs_err"] = abs(float(r["dice"]) - float(ref["dice"]))
    out["note"] = ("bf16 storage vs the fp32 reference; the reference's own logits move by 3.7e-2 rel-L2 when only its INPUT is "
                   "rounded to bf16 (MaxPool-argmax -> MaxUnpool flips; tests/golden/headline_gan_step_2x512.npz: sens_input_bf16)")
    del g, d, og, od
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

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
    from ssunet_gan_b200 import _lib, batchnorm, metrics, models_seg_gan, ops, optim, replicate, train_step

    ssg.set_compute_dtype(torch.bfloat16)
    ssg.set_conv_impl(args.conv)
    size = args.size or cfg["size"]
    cin = cfg["cin"]
    batch = args.batch or (cfg["batch1"] if world == 1 else cfg["batchn"])

    # ---- CPU reference leg + parity gate (rank 0 at N = 1, BEFORE anything is timed on the GPU) ----
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        cb, cs = cpu_sample_of(args, cfg)
        cpu = cpu_reference_steps(cb, cs, cin, steps=2, warmup=0)
        parity = parity_gate(cpu, cin, cb, cs)

    torch.manual_seed(41)          # train_seg_gan.py:35-36; G then D
    g = models_seg_gan.Generator({"arch": "UNet_R_SS_v2", "num_classes": 3, "input_channels": cin, "deep_supervision": False}).cuda().train()
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
        x, t = synthetic_batch(batch, cin, size, size, seed=1234 + rank + 97 * i)
        host.append((x.pin_memory(), t.pin_memory()))
    dev = [(x.cuda(), t.cuda()) for x, t in host]

    graphed = None
    if not args.eager:
        # the whole G+D iteration (fwd, losses, metric kernels, bwd, SyncBN exchanges, NCCL all-reduces, clamp+Adam, operand
        # re-packing) captured once and replayed
        graphed = train_step.GraphedGanStep(g, d, og, od, (batch, cin, size, size), with_metrics=True)

    def step_dev(i):
        x, t = dev[i % n_sets]
        if graphed is not None:
            return graphed(x, t)
        return train_step.gan_train_step(g, d, og, od, x, t, with_metrics="device")

    # end-to-end leg: every step copies ITS inputs from pinned host memory and reads loss, IoU and Dice back.  The copy of step
    # i+1 is issued on a side stream while step i computes (double-buffered device staging), as a data loader would.
    copy_stream = torch.cuda.Stream()
    staging = [(torch.empty_like(dev[0][0]), torch.empty_like(dev[0][1])) for _ in range(2)]
    staged_ev = [None, None]
    d2h_bytes = [4]

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
            r = train_step.gan_train_step(g, d, og, od, x, t, with_metrics="device")
        loss = float(r["loss"])            # D2H reads of the step's results (sync): loss, IoU counts, Dice leaf sums
        train_step.finish_metrics(r)
        counts, leaf, _n = r["metric_parts"]
        d2h_bytes[0] = 4 + counts.numel() * 8 + leaf.numel() * 4
        return loss, r["iou"], r["dice"]

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

    for i in range(max(3, args.warmup)):
        step_dev(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms = timed(step_dev, args.steps)
    launches = graphed.launches_per_step * args.steps if graphed is not None else (_lib.launch_count - l0)
    last_metrics = [None]

    def e2e_fn(i):
        last_metrics[0] = step_e2e(i, i == args.steps - 1)

    ms_e2e = timed(e2e_fn, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        ops.PeerStatReducer.check_all()          # a SyncBN exchange that timed out would have invalidated the run

    # roofline leg: the dominant kernels (tcgen05 convolutions) timed one by one with CUDA events on the launching stream,
    # live, over eager steps of the same workload (a captured graph cannot carry per-kernel events)
    prof_steps = 2
    PROF = ["ssg_conv2d_fwd_tc", "ssg_conv2d_dgrad_tc", "ssg_conv2d_dgrad_tc_acc", "ssg_conv2d_dgrad_tc_split", "ssg_conv2d_dgrad_tc_mask",
            "ssg_conv2d_wgrad_tc", "ssg_conv2d_wgrad_tc_acc"]
    train_step.gan_train_step(g, d, og, od, dev[0][0], dev[0][1], with_metrics="device")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.profile_reset(PROF)
    e0.record()
    for i in range(prof_steps):
        x, t = dev[i % n_sets]
        train_step.gan_train_step(g, d, og, od, x, t, with_metrics="device")
    e1.record()
    prof = _lib.profile_collect()
    ms_prof = e0.elapsed_time(e1)

    # SyncBN stat-reduce latency (BASELINE.json metric): the [sum x | sum x^2] fp64 all-reduce of one BN layer over NVLink,
    # timed alone on the compute stream (CUDA events, 50 back-to-back reduces after 10 warm-ups), smallest and largest layer
    stat_reduce_us = None
    if world > 1:
        peer = ops.PeerStatReducer.for_group(dist.group.WORLD)
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

    # BASELINE configs[4] (extra key, not the headline): inference-only segmentation through the eval-mode generator with
    # IoU + Dice from metrics.py's kernels.  `value`: batch resident in HBM; `e2e`: the batch copied from pinned host memory
    # every step (prefetched on a side stream) and both metric results read back.
    infer = None
    if world == 1 and args.infer_batch > 0 and args.config == "headline":
        g.eval()
        ib = args.infer_batch
        hx = torch.randn(ib, cin, size, size).pin_memory()
        ht = (torch.rand(ib, 3, size, size) > 0.5).float().pin_memory()
        xi, ti = hx.cuda(), ht.cuda()
        ti_m = ti[:, 1:].contiguous()

        def infer_step(_i, x=None):
            with torch.no_grad():
                lo = g(xi if x is None else x)
                om = lo[:, 1:].contiguous()
                return metrics.iou_score(om, ti_m), metrics.dice_coef(om, ti_m)

        for i in range(2):
            infer_step(i)
        ms_inf = timed(infer_step, 3)
        stage = [torch.empty_like(xi), torch.empty_like(xi)]
        evs = [None, None]

        def pre(i):
            with torch.cuda.stream(copy_stream):
                stage[i % 2].copy_(hx, non_blocking=True)
                ev = torch.cuda.Event(); ev.record(copy_stream)
            evs[i % 2] = ev

        def infer_e2e(i):
            if evs[i % 2] is None:
                pre(i)
            torch.cuda.current_stream().wait_event(evs[i % 2])
            evs[i % 2] = None
            if i < 2:
                pre(i + 1)
            return infer_step(i, stage[i % 2])

        ms_inf_e2e = timed(infer_e2e, 3)
        v = ib * 3 / (ms_inf / 1e3)
        peaks_tf = 1373.8
        infer = {"value": v, "unit": "img/s", "batch": ib, "ms_per_batch": ms_inf / 3,
                 "e2e": {"value": ib * 3 / (ms_inf_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": int(hx.numel() * 4), "d2h_bytes_per_step": 16 + 3 * 4 * ((ib * 2 * size * size + 127) // 128)},
                 "roofline": {"bound": "tensor", "achieved": v * G_FWD_GFLOP_512 * (size * size) / (512.0 * 512.0) / 1e3, "unit": "TFLOP/s",
                              "how": "417.25 GFLOP of convolutions per 512^2 image (SURVEY §8d) x images/s; whole forward incl. BN / SPADE / pooling passes"},
                 "what": "Generator.eval() forward + iou_score + dice_coef, batch %d x %d x %d x %d bf16 (BASELINE configs[4])" % (ib, cin, size, size)}
        del xi, ti, stage
        g.train()

    # Weak-scaling anchor (extra key, not the headline): BASELINE configs[1] runs batch 16 on one GPU but configs[2] runs batch 8 PER
    # GPU on 2/4/8 GPUs, so value(N) / (N * value(1)) compares different per-GPU work.  The same step at batch 8 on this one GPU is
    # the per-GPU work of the N > 1 runs: value(N) / (N * anchor) is the scaling efficiency at equal work per GPU.
    anchor = None
    if world == 1 and not args.batch and graphed is not None and batch > cfg["batchn"]:
        b8 = cfg["batchn"]
        xs = [(x[:b8].contiguous(), t[:b8].contiguous()) for x, t in dev]
        g8 = train_step.GraphedGanStep(g, d, og, od, (b8, cin, size, size), with_metrics=True)
        for i in range(2):
            g8(*xs[i % n_sets])
        ms8 = timed(lambda i: g8(*xs[i % n_sets]), args.steps)
        anchor = {"value": b8 * args.steps / (ms8 / 1e3), "unit": UNIT, "batch_per_gpu": b8,
                  "ms_per_step": ms8 / args.steps,
                  "what": "the same captured G+D step at batch %d on one GPU = the per-GPU work of BASELINE configs[2] (N = 2/4/8)" % b8}
        del g8

    imgs = world * batch * args.steps
    value = imgs / (ms / 1e3)
    e2e = imgs / (ms_e2e / 1e3)

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
                                        for k, v in sorted(prof.get("by_name", {}).items())}})
    # DRAM traffic of the dominant kernel family (conv_tc_halo_kernel: 36 % of the step's device time in the launch list,
    # profiles/r02_launches_summary.txt), per launch, from the committed `ncu --set full` capture of its level-0 instance
    # (profiles/r02_traffic.json, written by profiles/traffic_from_ncu.py from the .ncu-rep files)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["kernels"]
        top = ([k for k in tr if k["capture"].endswith("halo_fwd_l0_lean")] or [k for k in tr if k["capture"].endswith("halo_fwd_l0_2issuers")]
               or [k for k in tr if k["capture"].endswith("halo_wgrad_l0")])[0]
        roof["traffic"] = top["dram_bytes"]
        roof["traffic_detail"] = {"kernel": top["kernel"], "launch": top["what"], "dram_bytes": top["dram_bytes"],
                                  "algorithmic_bytes": top["algorithmic_bytes"], "source": "profiles/r02_traffic.json (" + top["capture"] + ".ncu-rep)",
                                  "other_kernels": {k["capture"].split("_", 1)[1]: k["traffic_over_algorithmic"] for k in tr}}
    except Exception:
        pass
    scale = (size * size) / (512.0 * 512.0)
    gflop_img = STEP_GFLOP_512 * scale + cfg["extra_gflop"]
    step_tf = gflop_img * 1e9 * world * batch * args.steps / (ms / 1e3) / 1e12
    line = {"metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (randn tiles, Bernoulli(0.5) masks; seed-41 default init)",
            "config": {"workload": workload_name(cfg, batch, size, world),
                       "global_batch": world * batch, "parallelism": "dp%d" % world, "conv_impl": args.conv,
                       "launch": "eager (one launch per kernel)" if graphed is None else "CUDA graph replay of the captured step",
                       "l2_policy": "inputs+activations per step (GBs) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(batch * (cin + 3) * size * size * 4), "d2h_bytes_per_step": int(d2h_bytes[0]),
                    "read_back": "loss (4 B) + IoU counts (16 B) + Dice pairwise leaf sums; last step: loss %.6f iou %.6f dice %.6f" % last_metrics[0]},
            "gpu_launches": launches, "step_tflops_required_work": step_tf, "step_frac_of_tensor_peak": step_tf / world / peak_tf,
            "roofline": roof, "clocks": clocks}
    if parity is not None:
        line["parity"] = parity
    if stat_reduce_us is not None:
        line["syncbn_stat_reduce_us"] = stat_reduce_us
    if infer is not None:
        infer["roofline"]["peak"] = peak_tf
        infer["roofline"]["frac"] = infer["roofline"]["achieved"] / peak_tf
        line["inference"] = infer
    if anchor is not None:
        line["weak_scaling_anchor"] = anchor
    if cpu is not None:
        cb, cs = cpu_sample_of(args, cfg)
        best = min(cpu["times"])
        v = cb * (cs * cs) / float(size * size) / best
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu["kind"],
                                "sample": "two full G+D steps (the reference's own modules: baseline/_ref)" if cpu["kind"] == "reference" else
                                          "two full G+D steps (oracle port: baseline/_ref not staged)",
                                "sample_shape": "batch %d x %d x %d x %d fp32, best of %s s, pixel-scaled to %d^2" % (
                                    cb, cin, cs, cs, "/".join("%.1f" % t for t in cpu["times"]), size)}
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
