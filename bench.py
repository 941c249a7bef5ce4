#!/usr/bin/env python
"""Benchmark of the SRCGAN G+D training step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One *step* = one ``SRCycleGAN.optimize_parameters`` (3 G_A + 3 G_B + 3 D_A + 3 D_B forwards, backward_G,
two D backwards, both Adam steps) over a batch of synthetic Sat2Aer-shaped patches
(real_B ~ U[0,1) 3x256x256, real_A = nearest 1/4 of it, random-init weights; no dataset/checkpoints
exist offline).  Workload at N=1 = BASELINE.json configs[1]: RDDBNet x4 G+D training, bf16, batch 64.
N>1 (torchrun, one rank per GPU): weak scaling, 64 patches per GPU, NCCL gradient all-reduce.

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks; L2 is not flushed between iterations because every iteration streams
tens of GB of activations through the 126 MB L2 (stated in config.l2).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "G+D train-step patches/sec (x4, 64->256)"
UNIT = "patches/s"
# algorithmic conv FLOPs of one patch through one step (fwd+dgrad+wgrad), BASELINE.md section 3
GFLOP_PER_PATCH = 3423.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="srcgan_b200", choices=["srcgan_b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="patches per GPU")
    ap.add_argument("--lr-size", type=int, default=64, help="LR patch edge (HR = 4x)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"tflops": float(p["bf16_tflops_sustained"]), "tflops_burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    except Exception:
        return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU every 200 ms while ``recording`` is set.

    NVML in-process (the library nvidia-smi itself queries): spawning an nvidia-smi process five times a second re-initialises
    NVML for every GPU of the box each time and was seen to stall a rank's launches for ~300 ms inside a 3-step timed window
    at N = 4.  Falls back to the nvidia-smi command line of B200_PROFILING.md when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.recording = index, [], False, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            pw = 0.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        act = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(mx), "%.1f" % pw, act(0x8), act(0x40), act(0x20), act(0x4)]   # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap

    def run(self):
        while not self.stop_flag:
            if self.recording:
                try:
                    if self.nvml is not None:
                        self.rows.append(self._sample_nvml())
                    else:
                        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                        self.rows.append([c.strip() for c in out.strip().split(",")])
                except Exception:
                    pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's step on the host cores
# --------------------------------------------------------------------------------------------

def cpu_reference_steps(batch: int, lr: int, steps: int, warmup: int):
    """-> (patches/s, cores, seconds per step).  fp32, all host threads."""
    import torch
    from oracle import srcgan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = O.CycleGANStepOracle(O.default_states(0))
    real_A, real_B = O.synthetic_batch(batch, lr=lr, scale=4, seed=1234)
    for _ in range(warmup):
        step.optimize_parameters(real_A, real_B)
    t0 = time.perf_counter()
    for _ in range(steps):
        step.optimize_parameters(real_A, real_B)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 1                      # bounded sample: ~9 s of CPU work per step at the full patch size
    steps, warmup = min(args.steps, 3), min(args.warmup, 1)
    v, cores, dt = cpu_reference_steps(batch, args.lr_size, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "RDDBNet x4 G+D training step (SRCycleGAN.optimize_parameters), 64x64->256x256 RGB",
                   "patches_per_step": batch, "note": "reference is pure Python/PyTorch: timed as the oracle port "
                   "(oracle/srcgan_oracle.py, pinned to the reference by tests/golden) on the host cores"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "batch %d x %d steps of the full-size step (fp32, torch CPU)" % (batch, steps)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from oracle import srcgan_oracle as O          # only for the shared deterministic weight init + cpu_baseline
    from srcgan_b200 import _lib, dist as sdist, nn as snn, ops, trainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the srcgan_b200 arm has no CPU fallback")
    local_rank = sdist.init_from_env()
    rank, world = sdist.rank(), sdist.world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    snn.set_precision(args.precision)
    peaks = load_peaks()

    opt = trainer.params()
    opt.device, opt.mode, opt.net = dev, "x4", "1"
    model = trainer.SRCycleGAN(opt)
    states = O.default_states(0)
    for name in ("G_A", "G_B", "D_A", "D_B"):
        getattr(model, "net" + name).load_state_dict(states[name], strict=True)
    sdist.make_data_parallel(model)

    import random
    random.seed(rank)
    B, lr = args.batch, args.lr_size
    g = torch.Generator().manual_seed(1234 + rank)
    host_B = torch.rand(B, 3, lr * 4, lr * 4, generator=g).pin_memory()
    real_B = host_B.to(dev, non_blocking=True)
    real_A = F.interpolate(real_B, scale_factor=0.25, mode="nearest")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        model.optimize_parameters(real_A, real_B)

    d2h_bytes = 0

    def step_e2e():
        nonlocal d2h_bytes
        rb = host_B.to(dev, non_blocking=True)                       # pinned host -> device, every step
        ra = F.interpolate(rb, scale_factor=0.25, mode="nearest")     # as train.py:381-382
        model.optimize_parameters(ra, rb)
        losses = model.current_losses()                               # device -> host read of the 9 losses
        d2h_bytes = 4 * len(losses)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                              # idle until the timed region begins
    for _ in range(args.warmup):
        step_resident()
    sampler.recording = True
    launches0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)                              # the reported value: no per-launch instrumentation
    launches = _lib.launch_count() - launches0
    # per-kernel timing for the roofline record: the same steps again with CUDA events around every convolution launch
    # (the event records cost ~8 % of a step, so this pass is separate and its wall time is only the share denominator)
    ops.timer.enabled = True
    ops.timer.reset()
    ms_instr = timed(step_resident, args.steps)
    ops.timer.enabled = False
    ksum = ops.timer.summary()
    ops.timer.reset()
    e2e = None
    if not args.no_e2e:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        e2e = {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": host_B.numel() * 4, "d2h_bytes_per_step": d2h_bytes}
    sampler.stop_flag = True

    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    value = B * world * args.steps / (ms * 1e-3)
    # dominant kernel family by device time
    roof = None
    if ksum:
        top = max(ksum.items(), key=lambda kv: kv[1]["ms"])
        name, d = top
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        conv_ms = sum(v["ms"] for v in ksum.values())
        roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops"], "traffic": None,
                "traffic_note": "launches of this kernel family span several layer shapes; per-shape DRAM bytes from ncu --set "
                                "full are in profiles/r1_ncu_sweep2.txt and profiles/r1_ncu_wgrad_stack.txt (64->32 @64x256x256: "
                                "563+248 MB measured vs 805 MB algorithmic; 192->64 @32x256x256: 897+247 MB vs 1074 MB)",
                "peak_source": peaks["source"],
                "launches_per_step": d["launches"] / args.steps, "avg_launch_ms": d["ms"] / d["launches"],
                "share_of_step": d["ms"] / ms_instr, "conv_share_of_step": conv_ms / ms_instr,
                "instrumented_ms_per_step": ms_instr / args.steps,
                # the dense-block layers (intensity 192-432 FLOP/B) sit near the machine's ridge: the same launches against
                # the measured copy bandwidth, from algorithmic bytes (input slice read once, output slice written once)
                "hbm": {"achieved": d["bytes"] / (d["ms"] * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": d["bytes"] / (d["ms"] * 1e-3) / 1e9 / peaks["hbm"]},
                "families": {k: {"ms_per_step": v["ms"] / args.steps, "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                                 "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                                 "launches_per_step": v["launches"] / args.steps} for k, v in sorted(ksum.items())}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "RDDBNet x4 G+D training step (SRCycleGAN.optimize_parameters), batch %d per GPU, "
                               "%dx%d->%dx%d RGB patches" % (B, lr, lr, lr * 4, lr * 4),
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "not flushed: each step streams >10 GB of activations per GPU through the 126 MB L2",
                   "algorithmic_gflop_per_patch": GFLOP_PER_PATCH * (lr / 64.0) ** 2,
                   "step_tflops": value * GFLOP_PER_PATCH * (lr / 64.0) ** 2 / 1e3 / world},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, dt = cpu_reference_steps(1, lr, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "1 step at batch 1 of the same full-size step, fp32, oracle port of the "
                                          "reference (pure PyTorch) on the host cores, %.1f s" % dt}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
