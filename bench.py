#!/usr/bin/env python
"""bench.py -- FlowNetS training-step throughput on synthetic FlyingChairs-shaped 384x512 pairs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math bf16|tf32|fp32]

One "step" = one pass of the hot path over one batch: H2D-resident inputs -> pre-processing -> conv
tower -> refinement -> fused warp+loss -> backward -> (gradient all-reduce) -> Adam.
N>1 is launched by torchrun (one rank per GPU, NCCL); per-GPU batch is fixed (weak scaling).
Rank 0 prints ONE JSON line (contract in the task statement / DESIGN.md section "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "FlyingChairs 384x512 training-step pairs/sec"
UNIT = "pairs/s"
H, W = 384, 512
PER_GPU_BATCH = 32          # BASELINE.json configs[1]: FlowNetS training, batch=32, 1xB200
TRAIN_GFLOP_PER_PAIR = None  # computed from the layer list below


# ----------------------------------------------------------------------------- workload arithmetic
def layer_flops(B: int, model: str = "flownets"):
    """Algorithmic FLOPs per launch tag (2*M*N*K of the implicit GEMM), DESIGN.md 'kernels'."""
    from deepof_b200.flownet import TOWER, REFINE
    fl = {}
    ih, iw = H, W
    tower = TOWER
    if model == "flownetc":
        h8, w8 = H // 8, W // 8
        tower = [("conv4_1", 3, 2, 256, 512), ("conv4_2", 3, 1, 512, 512), ("conv5_1", 3, 2, 512, 512), ("conv5_2", 3, 1, 512, 512),
                 ("conv6_1", 3, 2, 512, 1024), ("conv6_2", 3, 1, 1024, 1024)]
        for nm, k, s, cin, cout, hh, ww in [("conv1", 7, 2, 3, 64, H, W), ("conv2", 5, 2, 64, 128, H // 2, W // 2), ("conv3", 5, 2, 128, 256, H // 4, W // 4)]:
            f = 2.0 * B * (hh // s) * (ww // s) * cout * k * k * cin
            for br in "ab":
                fl[f"conv_fwd:{nm}{br}"] = f
                fl[f"conv_wgrad:{nm}{br}"] = f
                if nm != "conv1":
                    fl[f"conv_dgrad:{nm}{br}"] = f
        for nm, k, cin, cout in [("conv_redir", 1, 256, 32), ("conv3_1", 3, 473, 256)]:
            f = 2.0 * B * h8 * w8 * cout * k * k * cin
            fl["conv_fwd:" + nm] = fl["conv_wgrad:" + nm] = fl["conv_dgrad:" + nm] = f
        fl["corr_fwd"] = 2.0 * B * h8 * w8 * 441 * 256
        fl["corr_bwd"] = 2.0 * fl["corr_fwd"]
        ih, iw = h8, w8
    for name, k, s, cin, cout in tower:
        oh, ow = -(-ih // s), -(-iw // s)
        f = 2.0 * B * oh * ow * cout * k * k * cin
        fl["conv_fwd:" + name] = f
        fl["conv_wgrad:" + name] = f
        if name != "conv1":
            fl["conv_dgrad:" + name] = f
        ih, iw = oh, ow
    for s, cfeat, up, upc, _uppr, _skip in REFINE:
        hs, ws = H >> s, W >> s
        f = 2.0 * B * hs * ws * cfeat * 16 * upc          # 4x4 taps, every (small pixel, tap) pair used once
        fl["deconv_fwd:" + up] = f
        fl["deconv_dgrad:" + up] = f
        fl["deconv_wgrad:" + up] = f
    for s, c in {6: 1024, 5: 1026, 4: 770, 3: 386, 2: 194, 1: 98}.items():
        f = 2.0 * B * (H >> s) * (W >> s) * c * 9 * 2
        fl[f"head_fwd:pr{s}"] = f
        fl[f"head_dgrad:pr{s}"] = f
        fl[f"head_wgrad:pr{s}"] = f
    return fl


def layer_bytes(B: int, n_param_floats: int, lean: bool = False):
    """Algorithmic HBM bytes for the bandwidth-bound launches.  lean bf16 engine: the heads read the bf16 feature maps
    (2 B per channel) and move the 20-float Z map / the 64-column bf16 D9 map once each way."""
    by = {}
    px = sum((H >> s) * (W >> s) for s in range(1, 7)) * B
    by["warp_loss"] = px * (44 + 8) - (px - B * (H >> 1) * (W >> 1)) * 12      # recon written for scale 1 only
    by["adam"] = n_param_floats * 28
    for s, c in {6: 1024, 5: 1026, 4: 770, 3: 386, 2: 194, 1: 98}.items():
        n = B * (H >> s) * (W >> s)
        if lean:
            by[f"head_fwd:pr{s}"] = n * (c * 2 + 80)
            by[f"head_tapsum:pr{s}"] = n * (80 + 8)
            by[f"head_dpr9:pr{s}"] = n * (8 + 36)
            by[f"head_wgrad:pr{s}"] = n * (c * 2 + 128)
        else:
            by[f"head_fwd:pr{s}"] = n * (c * 4 + 8)
            by[f"head_dgrad:pr{s}"] = n * (c * 4 * (1 if s == 1 else 2) + 8)
            by[f"head_wgrad:pr{s}"] = n * (c * 4 + 8)
    return by


# ----------------------------------------------------------------------------- helpers
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.strip()]
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                r = [c.strip() for c in r]
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range seen
        thr = min(pw) + 0.5 * (max(pw) - min(pw))
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU / reference arm
def _host_cores() -> int:
    """Physical cores in this process's affinity mask, capped by the cgroup v2 CPU quota (oversubscribing OpenMP threads is far slower)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cpus = list(range(os.cpu_count() or 1))
    groups = set()
    for c in cpus:
        try:
            groups.add(open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list").read().strip())
        except OSError:
            groups.add(str(c))
    n = max(1, len(groups))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return n


def cpu_step_rate(sample_pairs: int, steps: int, warmup: int):
    """Times the oracle (reference-semantics CPU restatement, torch fp32) training step on the host cores.

    This is the one place bench.py executes oracle/: as the reported CPU baseline / the reference arm.
    The reference's own TF-0.1x graph cannot run here (no tensorflow, no python2 -- SURVEY.md 0.3)."""
    import torch
    from oracle import flownet_s as ofs, adam as oadam
    from deepof_b200.synth import make_pairs
    # the physical cores this process may run on, capped by the cgroup CPU quota: torchrun exports OMP_NUM_THREADS=1 (would make this arm
    # single-threaded) and torch's own default ignores the quota (64 threads on a 16-CPU quota on the bench boxes)
    torch.set_num_threads(_host_cores())
    threads = torch.get_num_threads()
    src, tgt, _ = make_pairs(sample_pairs, H, W, seed=0)
    params = ofs.init_params(1)
    opt = oadam.TFAdam(params)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _tot, grads, *_ = ofs.loss_and_grads(params, src, tgt)
        opt.step(grads, 1.6e-5)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return dict(value=sample_pairs / mean, unit=UNIT, cores=threads, kind="port",
                sample=f"{steps} training steps (fwd+bwd+TF-Adam) of {sample_pairs} pairs at {H}x{W}, torch-CPU fp32 oracle, "
                       f"{threads} threads, mean {mean * 1e3:.0f} ms/step"), mean


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 2
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 2))
    cb, mean = cpu_step_rate(sample, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"FlowNetS training, synthetic FlyingChairs {H}x{W}, batch={PER_GPU_BATCH} per GPU "
                                   f"(reference arm: bounded sample of {sample} pairs per step on the host CPU)",
                       "global_batch": PER_GPU_BATCH * args.gpus},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference-semantics CPU restatement (PyTorch fp32 oracle); the TF-0.1x reference cannot be executed in this image"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from deepof_b200 import _lib
    from deepof_b200.flyingChairsTrain import TrainStep, WEIGHT_L
    from deepof_b200.synth import make_pairs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    lib = _lib.load()
    step = TrainStep(B, (H, W), device=dev, math_mode=args.math, distributed=world > 1, tc_wgrad=(args.math != "fp32"),
                     model=args.model, variant=args.variant)
    eng = step.engine
    # two different synthetic batches per rank, alternated (working set per step ~3 GB >> 126 MB L2)
    batches = []
    for j in range(2):
        s, t, _ = make_pairs(B, H, W, seed=1000 * rank + j)
        batches.append((s.pin_memory(), t.pin_memory(), s.to(dev), t.to(dev)))
    lr = 1.6e-5

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        _, _, s, t = batches[i % 2]
        eng.train_step(s, t, WEIGHT_L, lr, allreduce=step.reducer)

    def e2e_step(i):
        s, t, _, _ = batches[i % 2]
        step.run({"source_img": s, "target_img": t, "loss_weight": WEIGHT_L, "learning_rate": lr})
        return step.last_loss(lag=1)       # D2H read of a step's loss every step (the previous step's: no device stall)

    # ---- device-resident timing ("value") ----
    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    lib.dofb_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        device_step(i)
    e1.record()
    barrier()
    launches = int(lib.dofb_launch_count())
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    loss_dev = float(eng.total_loss().item())

    # ---- end-to-end timing through the reference-facing API ("e2e") ----
    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    for i in range(args.steps):
        last = e2e_step(i)
    last = step.last_loss(lag=0)           # drain: the final step's loss is read inside the timed region as well
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- max over ranks ----
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    ddp_sync = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # replicas must hold bit-identical parameters after every step (same broadcast start, same reduced gradients)
        chk = eng.theta.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ddp_sync = bool((lo == hi).item())
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- per-launch timing for the roofline (rank 0, a few instrumented steps, CUDA events per launch) ----
    roof, breakdown = None, None
    if rank == 0:
        eng.profile = []
        psteps = min(args.steps, 3)
        for i in range(psteps):
            # local steps only (no collective: the other ranks are already done)
            eng.train_step(batches[i % 2][2], batches[i % 2][3], WEIGHT_L, lr, allreduce=None)
        torch.cuda.synchronize()
        per = {}
        for tag, a, b in eng.profile:
            per.setdefault(tag, []).append(a.elapsed_time(b))
        eng.profile = None
        avg = {k: sum(v) / len(v) * (len(v) / psteps) for k, v in per.items()}     # ms per step per tag
        fl = layer_flops(B, args.model)
        by = layer_bytes(B, eng.arena.n_true, getattr(eng, 'lean', False))
        peaks = load_peaks()
        classes = {}
        for tag, tms in avg.items():
            cls = tag.split(":")[0]
            c = classes.setdefault(cls, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            c["ms"] += tms
            c["flops"] += fl.get(tag, 0.0)
            c["bytes"] += by.get(tag, 0.0)
            c["launches"] += len(per[tag]) // psteps
        total_ms = sum(c["ms"] for c in classes.values())
        gemm = [k for k in classes if k.startswith("conv_") or k.startswith("deconv_")]   # (corr_* reported separately)
        gemm_ms = sum(classes[k]["ms"] for k in gemm)
        gemm_fl = sum(classes[k]["flops"] for k in gemm)
        gemm_n = sum(classes[k]["launches"] for k in gemm)
        # DRAM bytes per launch of the tcgen05 family, from the committed ncu capture of the same workload (profiles/r01_tc_traffic.json:
        # `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over the 44 tc_* launches of one step); only valid for that workload
        tc_traffic = None
        tr = ROOT / "profiles" / "r01_tc_traffic.json"
        if tr.exists() and args.math == "bf16" and args.model == "flownets" and B == PER_GPU_BATCH:
            tc_traffic = float(json.loads(tr.read_text())["dram_bytes_per_launch"])
        tf32 = args.math == "tf32"
        peak = peaks["bf16_sustained"] * (0.5 if tf32 else 1.0)        # bf16 math (and the fp32 SIMT path) are divided by the bf16 peak
        ach = gemm_fl / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "implicit-GEMM conv family (fwd/dgrad/wgrad, conv + transposed conv), "
                                             + {"tf32": "tcgen05 kind::tf32", "bf16": "tcgen05 kind::f16 (bf16 operands, fp32 accumulate)",
                                                "fp32": "SIMT fp32 FFMA (parity-grade path)"}[args.math],
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})" + (" x0.5 for tf32" if tf32 else ""),
                "traffic": tc_traffic, "traffic_source": "profiles/r01_tc_traffic.json (ncu dram bytes, average per tc_* launch)" if tc_traffic else None,
                "share_of_step": gemm_ms / total_ms, "launches_per_step": gemm_n,
                "avg_launch_ms": gemm_ms / max(gemm_n, 1)}
        breakdown = []
        for k, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
            e = {"class": k, "ms_per_step": round(c["ms"], 4), "share": round(c["ms"] / total_ms, 4), "launches": c["launches"]}
            if c["flops"]:
                e["tflops"] = round(c["flops"] / (c["ms"] * 1e-3) / 1e12, 3)
            if c["bytes"]:
                e["gbs"] = round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1)
                e["hbm_frac"] = round(e["gbs"] / peaks["hbm"], 4)
            breakdown.append(e)
        out_dir = ROOT / "gpurun_out"
        out_dir.mkdir(exist_ok=True)
        (out_dir / f"bench_layers_{args.model}_{args.math}_n{world}.json").write_text(json.dumps(
            {"per_tag_ms": {k: round(v, 5) for k, v in sorted(avg.items(), key=lambda kv: -kv[1])}, "classes": breakdown}, indent=1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- accuracy of the benched math mode on a held-out synthetic batch (north_star: EPE within 1e-3 of the fp32 forward) ----
    accuracy = None
    if not args.no_accuracy:
        from deepof_b200.flownet import FlowNetS, FlowNetC
        from deepof_b200 import ops as _ops
        cls = FlowNetS if args.model == "flownets" else FlowNetC
        hb = 4
        hs, ht, hgt = make_pairs(hb, H, W, seed=1234)
        hs, ht, hgt = hs.to(dev), ht.to(dev), hgt.to(dev)
        flows = {}
        for mode in ("fp32", args.math):
            if mode in flows:
                continue
            e = cls(hb, H, W, device=dev, math_mode=mode, seed=1, variant=args.variant, tc_wgrad=(mode != "fp32"))
            e.forward(hs, ht, with_grad=False)
            flows[mode] = (e.pr[1] * 10.0).clone()
            del e
        def epe(flow_s1):      # evaluation recipe of flyingChairsTrain.py:264-266,294-296: x2, clip, bilinear resize to HxW, AEE
            f = torch.clamp(flow_s1 * 2.0, -300.0, 250.0).permute(0, 3, 1, 2)
            f = torch.nn.functional.interpolate(f, size=(H, W), mode="bilinear", align_corners=False).permute(0, 2, 3, 1).contiguous()
            out = torch.zeros(1, dtype=torch.float64, device=dev)
            _ops.epe_sum(f, hgt.contiguous(), out)
            return out.item() / (hb * H * W)
        e32, em = epe(flows["fp32"]), epe(flows[args.math])
        accuracy = {"held_out_pairs": hb, "epe_fp32_path": e32, "epe_benched_path": em, "abs_epe_delta": abs(em - e32),
                    "flow_l1_mean_px": float((flows[args.math] - flows["fp32"]).abs().mean()),
                    "flow_l1_max_px": float((flows[args.math] - flows["fp32"]).abs().max()), "tolerance": 1e-3,
                    "note": "fp32 path == CPU oracle to 5e-7 px (tests/test_gpu_flownet.py)"}
        torch.cuda.empty_cache()
    # ---- the same step on the fp32 SIMT kernels (bit-auditable path), a few steps, for the record ----
    fp32_ref = None
    if world == 1 and args.math != "fp32" and not args.no_accuracy:
        del step, eng
        torch.cuda.empty_cache()
        s32 = TrainStep(B, (H, W), device=dev, math_mode="fp32", model=args.model, variant=args.variant)
        for i in range(2):
            s32.engine.train_step(batches[i % 2][2], batches[i % 2][3], WEIGHT_L, lr)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for i in range(3):
            s32.engine.train_step(batches[i % 2][2], batches[i % 2][3], WEIGHT_L, lr)
        g1.record()
        torch.cuda.synchronize()
        fp32_ref = {"math": "fp32 (SIMT FFMA kernels)", "ms_per_step": g0.elapsed_time(g1) / 3, "value": B * 3 / (g0.elapsed_time(g1) * 1e-3), "unit": UNIT}
        del s32
        torch.cuda.empty_cache()
    # ---- CPU baseline on the host cores (N=1 only, bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu, _ = cpu_step_rate(2, 5, 1)
    gb = B * world
    value = gb * args.steps / (ms * 1e-3)
    e2e = gb * args.steps / (ms_e2e * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.math], "data": "synthetic",
            "config": {"workload": f"{'FlowNetS' if args.model == 'flownets' else 'FlowNetC (correlation cost-volume)'} training "
                                   f"(fwd+bwd+Adam), synthetic FlyingChairs {H}x{W}, batch={B} per GPU",
                       "global_batch": gb, "parallelism": f"dp{world}",
                       "loss_variant": "A (flyingChairsWrapFlow.loss_interp)" if args.variant == "A" else "B (flyingChairsWrapFlow_vgg / version1 warpflow)",
                       "math": args.math, "l2": "inputs+activations per step (~3 GB) exceed the 126 MB L2; two alternating batches"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 2 * B * H * W * 3 * 4, "d2h_bytes_per_step": 4,
                    "api": "deepof_b200.flyingChairsTrain.TrainStep.run(feed_dict) + last_loss(lag=1): pinned-host inputs, "
                           "H2D on a copy stream double-buffered against the previous step, loss D2H read one step late"},
            "gpu_launches": launches, "ddp_replicas_in_sync": ddp_sync,
            "roofline": roof, "cpu_baseline": cpu, "accuracy": accuracy, "fp32_math_path": fp32_ref, "kernel_classes": breakdown,
            "loss_after": loss_dev, "loss_after_e2e": last}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default=os.environ.get("DEEPOF_MATH", "bf16"), choices=["fp32", "tf32", "bf16"],
                    help="bf16: tcgen05 kind::f16 on bf16 activation shadows + bf16 packed weights, fp32 accumulate/epilogue/master weights; tf32: tcgen05 tensor-core convolutions (TF32 operands, fp32 accumulate; EPE within 1e-3 of the fp32 path, "
                         "checked in this run); fp32: SIMT FFMA parity path")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (BASELINE config: 32)")
    ap.add_argument("--model", default="flownets", choices=["flownets", "flownetc"], help="flownets = BASELINE configs[1]; flownetc = configs[2]")
    ap.add_argument("--variant", default="A", choices=["A", "B"], help="loss_interp variant (A: flyingChairsWrapFlow, B: _vgg/version1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-accuracy", action="store_true", help="skip the held-out EPE check")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: re-launch under torchrun so that `python bench.py --gpus N` works stand-alone
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
