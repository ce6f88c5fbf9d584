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

# BASELINE.json configs[1..4].  "weak": fixed per-GPU batch; "strong": fixed GLOBAL batch split over the ranks.
CONFIGS = {
    "flownets32": dict(baseline_config=1, model="flownets", hw=(384, 512), batch=32, scaling="weak", variant="A",
                       what="FlowNetS training (fwd+bwd+Adam), synthetic FlyingChairs 384x512"),
    "flownetc32": dict(baseline_config=2, model="flownetc", hw=(384, 512), batch=32, scaling="weak", variant="A",
                       what="FlowNetC (correlation cost-volume) training, synthetic FlyingChairs 384x512"),
    "guided64x8": dict(baseline_config=3, model="flownetc", hw=(384, 512), batch=64, scaling="strong", variant="B",
                       what="FlowNetC guided: warp + photometric loss variant B (flyingChairsWrapFlow_vgg / version1 warpflow), global batch 64"),
    "sintel16x8": dict(baseline_config=4, model="flownets", hw=(448, 1024), batch=16, scaling="strong", variant="A", sintel=True,
                       what="Sintel fine-tune shape 448x1024 (436 padded), sintelTrain.py hyper-parameters, global batch 16"),
}


def config_kwargs(cfg):
    """Engine keyword arguments of a config (reference constants: sintelWrapFlow.py:773, sintelTrain.py:50-53,180)."""
    if cfg.get("sintel"):
        from deepof_b200.flownet import SINTEL_MEAN
        return dict(mean=SINTEL_MEAN, hyper=dict(epsilon=1e-4, alpha_c=0.3, alpha_s=0.3, lambda_smooth=0.0)), [16, 8, 4, 4, 2, 1]
    return {}, [16, 8, 4, 2, 1, 1]


# ----------------------------------------------------------------------------- workload arithmetic
def layer_flops(B: int, model: str = "flownets", H: int = H, W: int = W):
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


def layer_bytes(B: int, n_param_floats: int, lean: bool = False, H: int = H, W: int = W):
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
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML in a background thread (first
    sample within a millisecond of start(); the timed region of K = 10 steps is only ~60 ms, shorter than nvidia-smi's start-up);
    falls back to `nvidia-smi -lms` when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.004

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.p = self.f = self.thread = None
        self.rows = []          # (sm_mhz, max_mhz, power_w, reasons bitmask)
        self._stop = False
        self.source = None

    def _nvml_loop(self, nv, h, mx):
        while not self._stop:
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), mx, nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            time.sleep(self.PERIOD_S)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it lists indices
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._nv = nv
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h, mx), daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi"
        except Exception:
            self.p = None

    def stop(self):
        sm, mx, pw, reasons = [], [], [], set()
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self._nv
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            for s_, m_, p_, r_ in self.rows:
                sm.append(s_); mx.append(m_); pw.append(p_)
                for name, bit in bits.items():
                    if r_ & bit:
                        reasons.add(name)
        elif self.p is not None:
            time.sleep(0.15)
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            for r in [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.strip()]:
                try:
                    r = [c.strip() for c in r]
                    sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                except Exception:
                    continue
            os.unlink(self.f.name)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML binding and no nvidia-smi"]}
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range seen
        thr = min(pw) + 0.5 * (max(pw) - min(pw))
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons), "source": self.source}


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
    warm = max(1, min(args.warmup, 3))
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
def _barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(cfg_name, math_mode, per_gpu_batch, steps, warmup, rank, local_rank, world, want_clocks=False, want_profile=False):
    """One training configuration: W warm-up + K timed device-resident steps (`value`), then the same through TrainStep.run with pinned-host
    inputs (`e2e`); times are CUDA events, max over ranks.  Returns a dict (rank 0: incl. the per-launch roofline breakdown)."""
    import torch
    import torch.distributed as dist
    from deepof_b200 import _lib
    from deepof_b200.flyingChairsTrain import TrainStep
    from deepof_b200.synth import make_pairs
    cfg = CONFIGS[cfg_name]
    Hc, Wc = cfg["hw"]
    kw, weights = config_kwargs(cfg)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    B = per_gpu_batch
    step = TrainStep(B, (Hc, Wc), device=dev, math_mode=math_mode, distributed=world > 1, tc_wgrad=(math_mode != "fp32"),
                     model=cfg["model"], variant=cfg["variant"], **kw)
    eng = step.engine
    # two different synthetic batches per rank, alternated (working set per step >> 126 MB L2)
    batches = []
    for j in range(2):
        s_, t_, _ = make_pairs(B, Hc, Wc, seed=1000 * rank + j)
        # 8-bit images, as the reference's loader hands them to feed_dict (cv2.imread / cv2.resize arrays, flyingChairsLoader.py:64-80): the
        # end-to-end arm feeds the pinned uint8 arrays, the device-resident arm their float32 casts -- the same numbers
        s8, t8 = s_.round().clamp_(0, 255).to(torch.uint8), t_.round().clamp_(0, 255).to(torch.uint8)
        batches.append((s8.pin_memory(), t8.pin_memory(), s8.float().to(dev), t8.float().to(dev)))
    lr = 1.6e-5

    def device_step(i):
        _, _, s_, t_ = batches[i % 2]
        eng.train_step(s_, t_, weights, lr, allreduce=step.reducer)

    def e2e_step(i):
        s_, t_, _, _ = batches[i % 2]
        step.run({"source_img": s_, "target_img": t_, "loss_weight": weights, "learning_rate": lr})
        return step.last_loss(lag=1)       # D2H read of a step's loss every step (the previous step's: no device stall)

    for i in range(warmup):
        device_step(i)
    _barrier(world)
    sampler = ClockSampler(local_rank) if (rank == 0 and want_clocks) else None
    if sampler:
        sampler.start()
    lib.dofb_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(world)
    e0.record()
    for i in range(steps):
        device_step(i)
    e1.record()
    _barrier(world)
    launches = int(lib.dofb_launch_count())
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    loss_dev = float(eng.total_loss().item())
    # ---- end-to-end through the reference-facing API ----
    for i in range(min(warmup, 3)):
        e2e_step(i)
    _barrier(world)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(steps):
        e2e_step(i)
    last = step.last_loss(lag=0)           # drain: the final step's loss is read inside the timed region as well
    f1.record()
    _barrier(world)
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    ddp_sync = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # replicas must hold bit-identical parameters after every step: compare a position-weighted checksum AND the extremes
        th = eng.theta.double()
        idx = torch.arange(th.numel(), device=dev, dtype=torch.float64)
        chk = torch.stack([th.sum(), (th * (1.0 + idx / th.numel())).sum(), th.abs().max()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ddp_sync = bool((lo == hi).all().item())
    ms, ms_e2e = float(t[0]), float(t[1])
    gb = B * world
    out = {"config": cfg_name, "math": math_mode, "per_gpu_batch": B, "global_batch": gb, "hw": [Hc, Wc],
           "value": gb * steps / (ms * 1e-3), "ms_per_step": ms / steps, "e2e_value": gb * steps / (ms_e2e * 1e-3),
           "e2e_ms_per_step": ms_e2e / steps, "gpu_launches": launches, "ddp_replicas_in_sync": ddp_sync, "clocks": clocks,
           "loss_after": loss_dev, "loss_after_e2e": last, "h2d_bytes_per_step": batches[0][0].numel() * batches[0][0].element_size() * 2, "lean": bool(getattr(eng, "lean", False)),
           "n_params": eng.arena.n_true}
    # ---- per-launch timing for the roofline (rank 0, a few instrumented LOCAL steps, CUDA events per launch) ----
    if rank == 0 and want_profile:
        eng.profile = []
        psteps = min(steps, 3)
        for i in range(psteps):
            eng.train_step(batches[i % 2][2], batches[i % 2][3], weights, lr, allreduce=None)
        torch.cuda.synchronize()
        per = {}
        for tag, a_, b_ in eng.profile:
            per.setdefault(tag, []).append(a_.elapsed_time(b_))
        eng.profile = None
        out["per_tag_ms"] = {k: sum(v) / len(v) * (len(v) / psteps) for k, v in per.items()}
        out["per_tag_launches"] = {k: len(v) // psteps for k, v in per.items()}
    del step, eng, batches
    torch.cuda.empty_cache()
    return out


def roofline_from(m, B):
    """Conv-family roofline + per-class breakdown from the per-launch timings of measure(..., want_profile=True)."""
    cfg = CONFIGS[m["config"]]
    Hc, Wc = cfg["hw"]
    avg = m["per_tag_ms"]
    fl = layer_flops(B, cfg["model"], Hc, Wc)
    by = layer_bytes(B, m["n_params"], m["lean"], Hc, Wc)
    peaks = load_peaks()
    classes = {}
    for tag, tms in avg.items():
        cls = tag.split(":")[0]
        c = classes.setdefault(cls, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        c["ms"] += tms
        c["flops"] += fl.get(tag, 0.0)
        c["bytes"] += by.get(tag, 0.0)
        c["launches"] += m["per_tag_launches"][tag]
    total_ms = sum(c["ms"] for c in classes.values())
    gemm = [k for k in classes if k.startswith("conv_") or k.startswith("deconv_")]   # (corr_* reported separately)
    gemm_ms = sum(classes[k]["ms"] for k in gemm)
    gemm_fl = sum(classes[k]["flops"] for k in gemm)
    gemm_n = sum(classes[k]["launches"] for k in gemm)
    math_mode = m["math"]
    tf32 = math_mode == "tf32"
    peak = peaks["bf16_sustained"] * (0.5 if tf32 else 1.0)        # bf16 math (and the fp32 SIMT path) are divided by the bf16 peak
    ach = gemm_fl / (gemm_ms * 1e-3) / 1e12
    # DRAM bytes per launch of the tcgen05 family from the committed ncu capture of THIS build and workload, if there is one
    tc_traffic, tsrc = None, None
    tr = ROOT / "profiles" / f"r02_tc_traffic_{math_mode}.json"
    if tr.exists() and m["config"] == "flownets32" and B == PER_GPU_BATCH:
        d = json.loads(tr.read_text())
        tc_traffic, tsrc = float(d["dram_bytes_per_launch"]), f"profiles/{tr.name} (ncu dram__bytes_read+write, average per tc_* launch, build {d.get('build', '?')})"
    roof = {"bound": "tensor", "kernel": "implicit-GEMM conv family (fwd/dgrad/wgrad, conv + transposed conv), "
                                         + {"tf32": "tcgen05 kind::tf32", "bf16": "tcgen05 kind::f16 (bf16 operands, fp32 accumulate)",
                                            "fp32": "SIMT fp32 FFMA (parity-grade path)"}[math_mode],
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})" + (" x0.5 for tf32" if tf32 else ""),
            "traffic": tc_traffic, "traffic_source": tsrc,
            "share_of_step": gemm_ms / total_ms, "launches_per_step": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
            "whole_step_frac": sum(fl.values()) / (m["ms_per_step"] * 1e-3) / 1e12 / peak}
    breakdown = []
    for k, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"class": k, "ms_per_step": round(c["ms"], 4), "share": round(c["ms"] / total_ms, 4), "launches": c["launches"]}
        if c["flops"]:
            e["tflops"] = round(c["flops"] / (c["ms"] * 1e-3) / 1e12, 3)
        if c["bytes"]:
            e["gbs"] = round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1)
            e["hbm_frac"] = round(e["gbs"] / peaks["hbm"], 4)
        breakdown.append(e)
    return roof, breakdown


def accuracy_block(dev, modes, train_steps=200):
    """Oracle-anchored precision evidence (north_star: per-pixel flow error and multi-scale EPE against the reference-semantics fp32 CPU
    forward): 8 held-out synthetic pairs (seed 1234), all six scales, on the initial weights AND on weights after `train_steps`
    optimiser steps (non-trivial flows).  The CPU oracle runs ONLY here as the checker."""
    import torch
    from oracle import flownet_s as ofs
    from deepof_b200.flownet import FlowNetS, FLOW_SCALES
    from deepof_b200 import precision
    from deepof_b200.flyingChairsLoader import evaluate_aee
    from deepof_b200.synth import make_pairs
    hb = 8
    torch.set_num_threads(_host_cores())
    hs, ht, hgt = make_pairs(hb, H, W, seed=1234)
    # trained weights: the bf16 engine, 200 steps on 4 alternating synthetic batches at 10x the reference's learning rate
    teng = FlowNetS(hb, H, W, device=dev, math_mode="bf16", seed=1, tc_wgrad=True)
    tb = [tuple(x.to(dev) for x in make_pairs(hb, H, W, seed=500 + j)[:2]) for j in range(4)]
    for i in range(train_steps):
        teng.train_step(*tb[i % 4], lr=1.6e-4)
    trained = teng.export_params()
    del teng, tb
    torch.cuda.empty_cache()
    out = {"held_out_pairs": hb, "seed": 1234, "scales": 6, "trained_steps": train_steps, "tolerance": precision.TOLERANCE,
           "reference": "fp32 CPU oracle forward (oracle/flownet_s.py, reference-semantics restatement) on the same weights and inputs"}
    for which, params in (("init", ofs.init_params(1)), ("trained", trained)):
        with torch.no_grad():
            _l, flows_ref, _p, _t = ofs.forward(params, hs, ht)
        epe_ref = evaluate_aee(flows_ref[0].to(dev), hgt.to(dev))
        blk = {"oracle_flow_magnitude_px": [round(v, 5) for v in precision.flow_magnitude(flows_ref)], "epe_oracle": epe_ref, "modes": {}}
        for mode in modes:
            e = FlowNetS(hb, H, W, device=dev, math_mode=mode, seed=None, tc_wgrad=(mode != "fp32"))
            e.load_params(params)
            e.forward(hs.to(dev), ht.to(dev), with_grad=False)
            flows = [e.pr[s] * FLOW_SCALES[s] for s in range(1, 7)]
            st = precision.end_point_distance(flows, flows_ref)
            epe = evaluate_aee(flows[0].contiguous(), hgt.to(dev))
            blk["modes"][mode] = {"epd_mean_px": [round(m_, 7) for m_, _x in st], "epd_max_px": [round(x_, 6) for _m, x_ in st],
                                  "epe": epe, "abs_epe_delta": abs(epe - epe_ref),
                                  "within_tolerance": bool(precision.within(st, mode) and abs(epe - epe_ref) <= precision.TOLERANCE[mode]["epe_abs"])}
            del e
            torch.cuda.empty_cache()
        out[which] = blk
    return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL / the launcher print while the communicator comes up (the "NCCL version ..."
        # banner) is sent to stderr by pointing file descriptor 1 at it for the duration of the initialisation
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            t0 = torch.zeros(1, device=dev)
            dist.all_reduce(t0)                  # (communicator fully up before the descriptor is restored)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    cfg = CONFIGS[args.config]
    B = args.batch if args.batch else (cfg["batch"] if cfg["scaling"] == "weak" else max(1, cfg["batch"] // world))
    main_m = measure(args.config, args.math, B, args.steps, args.warmup, rank, local_rank, world, want_clocks=True, want_profile=True)
    # ---- the other tensor-core math mode on the same workload (the line carries both; `value` is args.math) ----
    modes = {}
    other_math = [] if args.no_extras else [m for m in ("tf32",) if m != args.math]
    for mm in other_math:
        r = measure(args.config, mm, B, max(3, min(args.steps, 5)), 3, rank, local_rank, world, want_profile=True)
        modes[mm] = r
    # ---- BASELINE configs[2..4] on this many GPUs (strong-scaling configs split their global batch over the ranks) ----
    others = {}
    if not args.no_extras:
        for name, c in CONFIGS.items():
            if name == args.config:
                continue
            if c["scaling"] == "strong" and c["batch"] % world:
                continue
            bo = c["batch"] if c["scaling"] == "weak" else c["batch"] // world
            try:
                others[name] = measure(name, args.math, bo, max(3, min(args.steps, 5)), 3, rank, local_rank, world)
            except Exception as ex:      # a secondary configuration must not take the headline down with it
                others[name] = {"config": name, "error": f"{type(ex).__name__}: {ex}"[:300]}
                torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    roof, breakdown = roofline_from(main_m, B)
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    (out_dir / f"bench_layers_{cfg['model']}_{args.math}_n{world}.json").write_text(json.dumps(
        {"per_tag_ms": {k: round(v, 5) for k, v in sorted(main_m["per_tag_ms"].items(), key=lambda kv: -kv[1])}, "classes": breakdown}, indent=1))

    def mode_entry(m):
        e = {"value": m["value"], "ms_per_step": m["ms_per_step"], "e2e": m["e2e_value"], "unit": UNIT, "gpu_launches": m["gpu_launches"]}
        if "per_tag_ms" in m:
            r, _ = roofline_from(m, m["per_gpu_batch"])
            e["roofline"] = {k: r[k] for k in ("achieved", "peak", "frac", "unit", "share_of_step", "whole_step_frac")}
        return e
    modes_out = {args.math: mode_entry(main_m)}
    for mm, r in modes.items():
        modes_out[mm] = mode_entry(r)
    others_out = {}
    for name, r in others.items():
        if "error" in r:
            others_out[name] = r
            continue
        c = CONFIGS[name]
        others_out[name] = {"baseline_config": c["baseline_config"], "workload": c["what"], "scaling": c["scaling"], "hw": r["hw"],
                            "per_gpu_batch": r["per_gpu_batch"], "global_batch": r["global_batch"], "value": r["value"], "unit": UNIT,
                            "ms_per_step": r["ms_per_step"], "e2e": r["e2e_value"], "gpu_launches": r["gpu_launches"],
                            "ddp_replicas_in_sync": r["ddp_replicas_in_sync"], "math": r["math"]}
    # ---- accuracy of the math modes against the CPU oracle (rank 0, N=1 runs only: it needs the host cores) ----
    accuracy = None
    if world == 1 and not args.no_accuracy and args.config == "flownets32":
        accuracy = accuracy_block(dev, ["fp32", "tf32", "bf16"])
    # ---- the same step on the fp32 SIMT kernels (bit-auditable path), a few steps, for the record ----
    fp32_ref = None
    if world == 1 and args.math != "fp32" and not args.no_accuracy and not args.no_extras:
        r = measure(args.config, "fp32", B, 3, 2, rank, local_rank, world)
        fp32_ref = {"math": "fp32 (SIMT FFMA kernels)", "ms_per_step": r["ms_per_step"], "value": r["value"], "unit": UNIT}
    # ---- CPU baseline on the host cores (N=1 only, bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu, _ = cpu_step_rate(2, 5, 1)
    Hc, Wc = cfg["hw"]
    line = {"metric": METRIC, "value": main_m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_m["ms_per_step"], "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": {"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.math], "data": "synthetic",
            "config": {"workload": f"{cfg['what']}, batch={B} per GPU (BASELINE.json configs[{cfg['baseline_config']}])",
                       "name": args.config, "global_batch": B * world, "parallelism": f"dp{world}",
                       "loss_variant": "A (flyingChairsWrapFlow.loss_interp)" if cfg["variant"] == "A" else "B (flyingChairsWrapFlow_vgg / version1 warpflow)",
                       "math": args.math, "schedule": "lean bf16 (bf16-only activations, tap-in-N flow heads)" if main_m["lean"] else "classic",
                       "l2": "inputs+activations per step (~2 GB) exceed the 126 MB L2; two alternating batches"},
            "clocks": main_m["clocks"],
            "e2e": {"value": main_m["e2e_value"], "unit": UNIT, "ms_per_step": main_m["e2e_ms_per_step"],
                    "h2d_bytes_per_step": main_m["h2d_bytes_per_step"], "d2h_bytes_per_step": 4,
                    "api": "deepof_b200.flyingChairsTrain.TrainStep.run(feed_dict) + last_loss(lag=1): pinned-host uint8 BGR images (the dtype "
                           "flyingChairsLoader.py:64-80 returns and flyingChairsTrain.py:178 feeds; cast on the device by the pre-processing kernel), "
                           "H2D on a copy stream double-buffered against the previous step, loss D2H read one step late"},
            "gpu_launches": main_m["gpu_launches"], "ddp_replicas_in_sync": main_m["ddp_replicas_in_sync"],
            "roofline": roof, "cpu_baseline": cpu, "modes": modes_out, "other_configs": others_out, "accuracy": accuracy,
            "fp32_math_path": fp32_ref, "kernel_classes": breakdown,
            "loss_after": main_m["loss_after"], "loss_after_e2e": main_m["loss_after_e2e"]}
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
                    help="bf16: tcgen05 kind::f16 on bf16 activations + bf16 packed weights, fp32 accumulate / master weights / loss / Adam "
                         "(lean schedule); tf32: tcgen05 kind::tf32 on the fp32 buffers; fp32: SIMT FFMA parity path.  The line always carries "
                         "the tf32 numbers as well (modes) and the oracle-anchored accuracy of all three")
    ap.add_argument("--config", default="flownets32", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: configs[1], the one the metric is quoted on)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: the configuration's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-accuracy", action="store_true", help="skip the oracle-anchored accuracy block")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary math mode and BASELINE configs[2..4]")
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
