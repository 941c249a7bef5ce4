#!/usr/bin/env python
"""Benchmark of the SRCGAN G+D training step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (baseline/_ref, else the oracle port)
    python bench.py --impl torch-gpu --batch 16              # the incumbent: reference modules on stock PyTorch/cuDNN, same GPU

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
    ap.add_argument("--impl", default="srcgan_b200", choices=["srcgan_b200", "reference", "torch-gpu"])
    ap.add_argument("--variant", default="", help="torch-gpu arm: comma list of fp32,tf32,bf16-autocast-channels_last")
    ap.add_argument("--workload", default="gd", choices=["gd", "cascade", "cascade_lab", "eval"],
                    help="gd = BASELINE configs[1] (default, the driver's line); cascade / cascade_lab / eval = configs[2..4]")
    ap.add_argument("--batch", type=int, default=None, help="units per GPU per step (default: 64 patches; eval: 16 tiles)")
    ap.add_argument("--tile-batch", type=int, default=8, help="eval workload: tiles per generator forward (1 = the reference's loop)")
    ap.add_argument("--overlap", action="store_true", help="N>1: launch each network's bucket all-reduce from inside backward "
                    "(default: in the optimizer step pre-hook; the overlap loses time, see srcgan_b200/dist.py)")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: flat cat / all-reduce / copy-back in the step pre-hook "
                    "instead of the bucket reducer")
    ap.add_argument("--lr-size", type=int, default=64, help="LR patch edge (HR = 4x)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--by-shape", action="store_true",
                    help="key the roofline record's families by kernel | call shape instead of by kernel family")
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

REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")


def reference_staged() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "train.py"))


def import_reference_train():
    """The reference's UNMODIFIED ``train`` module from baseline/_ref/src (scripts/stage_reference.py), with its own
    ``model`` / ``losses`` packages.  visdom / skimage (logging, image I/O) are stubbed; the class ``RDDBNetA`` that
    train.py:11 imports but the reference never defines is the documented shim, built from reference classes only:
    ``model.model.RDDBNet`` constructor + ``model.model.Decoder``, forward = the commented-out one at model.py:370-378."""
    import importlib
    import types
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    sys.modules.setdefault("visdom", types.SimpleNamespace(Visdom=lambda *a, **k: None))
    try:
        import skimage  # noqa: F401
    except Exception:
        sk = types.ModuleType("skimage")
        sk.io, sk.color = types.ModuleType("skimage.io"), types.ModuleType("skimage.color")
        sk.color.lab2rgb = sk.color.rgb2lab = sk.color.rgb2gray = None
        sk.io.imsave = sk.io.imread = None
        sys.modules.update({"skimage": sk, "skimage.io": sk.io, "skimage.color": sk.color})
    pkg = importlib.import_module("model")
    mm = importlib.import_module("model.model")
    if not hasattr(pkg, "RDDBNetA"):
        class RDDBNetA(mm.RDDBNet):
            def __init__(self, in_nc, out_nc, nf, nb, gc=32, mode="x2"):
                super().__init__(in_nc, out_nc, nf, nb, gc=gc, mode=mode)
                self.decode = mm.Decoder()

            def forward(self, x):
                fea = self.conv_first(x)
                fea = fea + self.trunk_conv(self.RRDB_trunk(fea))
                return self.conv_last(self.decode(fea))
        pkg.RDDBNetA = RDDBNetA
    for name in ("RDDBNetB", "NLayerDiscriminator", "SRDenseNetA", "SRDenseNetB"):
        if not hasattr(pkg, name):
            setattr(pkg, name, getattr(mm, name))
    return importlib.import_module("train")


def synthetic_patches(batch: int, lr: int, seed: int):
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    real_B = torch.rand(batch, 3, lr * 4, lr * 4, generator=g)
    return F.interpolate(real_B, scale_factor=0.25, mode="nearest"), real_B


def cpu_reference_steps(batch: int, lr: int, steps: int, warmup: int):
    """-> (patches/s, cores, seconds per step, kind).  fp32, all host threads.  kind "reference": the reference's own
    train.SRCycleGAN from baseline/_ref; "port": the oracle restatement when that tree is not staged."""
    import random
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    real_A, real_B = synthetic_patches(batch, lr, 1234)
    if reference_staged():
        train = import_reference_train()
        opt = train.params()
        opt.device, opt.mode, opt.net = torch.device("cpu"), "x4", "1"
        torch.manual_seed(0)
        random.seed(0)
        step, kind = train.SRCycleGAN(opt), "reference"
    else:
        from oracle import srcgan_oracle as O
        step, kind = O.CycleGANStepOracle(O.default_states(0)), "port"
    for _ in range(warmup):
        step.optimize_parameters(real_A, real_B)
    t0 = time.perf_counter()
    for _ in range(steps):
        step.optimize_parameters(real_A, real_B)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, cores, dt, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "gd":
        v, cores, dt, kind, sample = cpu_cascade(args.workload == "cascade_lab", 1) if args.workload != "eval" else cpu_eval()
        unit = "tiles/s" if args.workload == "eval" else UNIT
        print(json.dumps({"impl": "reference", "metric": {"cascade": "cascade train-step patches/sec (x4)",
                                                            "cascade_lab": "cascade train-step patches/sec (ConstLAB)",
                                                            "eval": "eval sweep tiles/sec (512x512, x4)"}[args.workload],
                          "value": v, "unit": unit, "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": args.workload, "units_per_step": 1},
                          "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    batch = 1                      # bounded sample: ~9 s of CPU work per step at the full patch size
    steps, warmup = min(args.steps, 3), min(args.warmup, 1)
    v, cores, dt, kind = cpu_reference_steps(batch, args.lr_size, steps, warmup)
    how = ("the reference's own unmodified train.SRCycleGAN (baseline/_ref/src, RDDBNetA = the documented shim)" if kind == "reference"
           else "the oracle port (oracle/srcgan_oracle.py, pinned to the reference by tests/golden)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "RDDBNet x4 G+D training step (SRCycleGAN.optimize_parameters), 64x64->256x256 RGB",
                   "patches_per_step": batch, "note": "reference is pure Python/PyTorch: timed as %s on the host cores" % how},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "batch %d x %d steps of the full-size step (fp32, torch CPU)" % (batch, steps)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# the incumbent: the same unmodified reference modules through stock PyTorch / cuDNN on the same B200
# ("the kernel to beat on the same box")
# --------------------------------------------------------------------------------------------

def run_torch_gpu(args):
    """Prints one JSON line per variant: fp32 with TF32 off (the reference as shipped), fp32 with TF32 on, and
    autocast(bf16) + channels_last + cudnn.benchmark (what a user gets from stock PyTorch with the usual switches)."""
    import random
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "torch-gpu", "unavailable": "no CUDA device"}))
        return
    dev = torch.device("cuda", 0)
    lr = args.lr_size
    if reference_staged():
        train, src = import_reference_train(), "reference modules (baseline/_ref/src) + RDDBNetA shim"
    else:
        train, src = None, "oracle port (baseline/_ref not staged)"
    variants = [("fp32", dict(tf32=False, autocast=False, cl=False)),
                ("tf32", dict(tf32=True, autocast=False, cl=False)),
                ("bf16-autocast-channels_last", dict(tf32=True, autocast=True, cl=True))]
    if args.variant:
        variants = [v for v in variants if v[0] in args.variant.split(",")]
    for name, cfg in variants:
        torch.backends.cuda.matmul.allow_tf32 = cfg["tf32"]
        torch.backends.cudnn.allow_tf32 = cfg["tf32"]
        torch.backends.cudnn.benchmark = True
        batch = args.batch or 16
        while True:
            try:
                torch.manual_seed(0)
                random.seed(0)
                if train is not None:
                    opt = train.params()
                    opt.device, opt.mode, opt.net = dev, "x4", "1"
                    model = train.SRCycleGAN(opt)
                    nets = [model.netG_A, model.netG_B, model.netD_A, model.netD_B]
                    if cfg["cl"]:
                        for n in nets:
                            n.to(memory_format=torch.channels_last)
                    step = model.optimize_parameters
                else:
                    from oracle import srcgan_oracle as O
                    st = {k: {kk: vv.to(dev) for kk, vv in sd.items()} for k, sd in O.default_states(0).items()}
                    model = O.CycleGANStepOracle(st)
                    step = model.optimize_parameters
                real_A, real_B = (t.to(dev) for t in synthetic_patches(batch, lr, 1234))
                if cfg["cl"]:
                    real_A, real_B = real_A.contiguous(memory_format=torch.channels_last), real_B.contiguous(memory_format=torch.channels_last)

                def one():
                    if cfg["autocast"]:
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            step(real_A, real_B)
                    else:
                        step(real_A, real_B)
                for _ in range(max(args.warmup, 2)):
                    one()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(args.steps):
                    one()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / args.steps
                peak_gb = torch.cuda.max_memory_allocated() / 2 ** 30
                break
            except torch.cuda.OutOfMemoryError:
                model = step = real_A = real_B = None
                import gc
                gc.collect()
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
                if batch == 1:
                    raise
                batch //= 2
        line = {"impl": "torch-gpu", "variant": name, "metric": METRIC, "value": batch / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                "steps": args.steps, "warmup": max(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True,
                "dtype": name, "data": "synthetic",
                "config": {"workload": "RDDBNet x4 G+D training step through stock PyTorch %s / cuDNN %s, batch %d (requested %d; "
                                       "halved on out-of-memory), %dx%d->%dx%d RGB" % (torch.__version__, torch.backends.cudnn.version(),
                                                                                     batch, args.batch or 16, lr, lr, lr * 4, lr * 4),
                           "source": src, "peak_memory_gb": peak_gb},
                "step_tflops": batch / (ms * 1e-3) * GFLOP_PER_PATCH * (lr / 64.0) ** 2 / 1e3}
        print(json.dumps(line), flush=True)
        model = step = real_A = real_B = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------

def cpu_cascade(lab: bool, batch: int):
    """-> (patches/s, cores, seconds, kind, sample): one CasSRC.optimize_parameters on the host cores (the reference's own
    trainCas*.py from baseline/_ref when staged, else the oracle port)."""
    import torch
    variant, up, sr_name = ("ConstLAB", 2, "SRDN") if lab else ("", 4, "RDDBNet")
    import time as _t
    torch.set_num_threads(os.cpu_count() or 1)
    ra = torch.rand(batch, 1, 256, 256)
    rb = torch.rand(batch, 3, 256, 256)
    kind = "port"
    if reference_staged():
        import importlib
        import_reference_train()                                  # stubs + the reference's model package on sys.path
        mod = importlib.import_module("trainCas" + variant)
        ropt = mod.params()
        ropt.device = torch.device("cpu")
        ropt.up, ropt.SRModel, ropt.CModel = up, sr_name, "ResDeconv"
        m = mod.CasSRC(ropt)
        m.init_log()
        step, kind = (lambda: m.optimize_parameters(ra, rb)), "reference"
    else:
        from oracle import srcgan_oracle as O
        if lab:
            o = O.CascadeStepOracle(O.init_srdn(1), O.init_resdeconv(2, 2), lambda s_, t_: O.srdn(s_, t_),
                                    lambda s_, t_: O.resdeconv(s_, t_), 2, "ConstLAB")
        else:
            o = O.CascadeStepOracle(O.init_rddbnet_pkg(1, 1, 1, 4), O.init_resdeconv(2, 3),
                                    lambda s_, t_: O.rddbnet_pkg(s_, t_, 4), lambda s_, t_: O.resdeconv(s_, t_), 4, "")
        step = lambda: o.optimize_parameters(ra, rb)
    t0 = _t.perf_counter()
    step()
    dt = _t.perf_counter() - t0
    return batch / dt, os.cpu_count() or 1, dt, kind, "1 step at batch %d, fp32, %s on the host cores, %.1f s" % (
        batch, "the reference's own trainCas%s.CasSRC (baseline/_ref)" % variant if kind == "reference" else "oracle port", dt)



def cpu_eval():
    """-> (tiles/s, cores, seconds, kind, sample): one 512x512 tile through the reference's RDDBNetB + metrics.py on the host."""
    import torch
    import torch.nn.functional as F
    import time as _t
    torch.set_num_threads(os.cpu_count() or 1)
    h = torch.rand(1, 3, 512, 512)
    l = F.interpolate(h, scale_factor=0.25, mode="nearest")
    kind = "port"
    if reference_staged():
        import importlib
        import_reference_train()
        mm, met = importlib.import_module("model.model"), importlib.import_module("metrics")
        rnet = mm.RDDBNetB(3, 3, 64, nb=3, mode="x4").eval()
        evs = [met.MSE(), met.PSNR(), met.AE(), met.SSIM()]

        def step():
            out = rnet(l)                                          # the reference builds the graph here too (no no_grad)
            return [float(torch.as_tensor(e(out.detach(), h)).mean()) for e in evs]
        kind = "reference"
    else:
        from oracle import srcgan_oracle as O
        sd = O.init_rddbnet_b(1)

        def step():
            out = O.rddbnet_b(sd, l, "x4")
            return [float(O.mse_loss(out, h)), float(O.psnr(out, h)), float(O.angular_error(out, h).mean()), float(O.ssim(out, h))]
    t0 = _t.perf_counter()
    step()
    dt = _t.perf_counter() - t0
    return 1.0 / dt, os.cpu_count() or 1, dt, kind, "1 tile (128x128 -> 512x512), fp32, %s on the host cores, %.1f s" % (
        "the reference's RDDBNetB + metrics.py (baseline/_ref)" if kind == "reference" else "oracle port", dt)



# ---- workloads (BASELINE.json configs; the default "gd" = configs[1] is the line the driver records) --------------------
# each builder returns a dict: name, unit, units_per_step (per rank), step_resident(), step_e2e() -> d2h bytes,
# h2d_bytes, gflop_per_unit (algorithmic, or None), cpu(batch) -> (units/s, seconds, kind, sample text), dp_nets/dp_opts

def wl_gd(args, dev, rank):
    """configs[1]: RDDBNet x4 G+D training step, 64x64 -> 256x256 RGB patches (src/train.py:325-340)."""
    import random
    import torch
    import torch.nn.functional as F
    from srcgan_b200 import trainer
    opt = trainer.params()
    opt.device, opt.mode, opt.net = dev, "x4", "1"
    torch.manual_seed(0)                            # random-init weights of the reference's distributions (package ctor)
    model = trainer.SRCycleGAN(opt)
    random.seed(rank)
    B, lr = args.batch, args.lr_size
    g = torch.Generator().manual_seed(1234 + rank)
    host_B = torch.rand(B, 3, lr * 4, lr * 4, generator=g).pin_memory()
    real_B = host_B.to(dev, non_blocking=True)
    real_A = F.interpolate(real_B, scale_factor=0.25, mode="nearest")

    def step_resident():
        model.optimize_parameters(real_A, real_B)

    feed = [None]

    def step_e2e():
        if feed[0] is None:       # the training loop's host -> device feed: every step's batch is copied from pinned host memory,
            import itertools      # one step ahead on a side stream (srcgan_b200/data.py), so the copy overlaps the previous step
            from srcgan_b200 import data
            feed[0] = data.DevicePrefetcher(itertools.repeat(host_B), dev)
        rb = next(feed[0])                                            # pinned host -> device, every step
        ra = F.interpolate(rb, scale_factor=0.25, mode="nearest")     # as train.py:381-382
        model.optimize_parameters(ra, rb)
        return 4 * len(model.current_losses())                        # device -> host read of the 9 losses

    def cpu(batch):
        v, cores, dt, kind = cpu_reference_steps(batch, lr, 1, 0)
        return v, cores, dt, kind, "1 step at batch %d of the same full-size step, fp32, %s on the host cores, %.1f s" % (
            batch, "the reference's own train.SRCycleGAN (baseline/_ref)" if kind == "reference" else "oracle port", dt)

    return dict(name="RDDBNet x4 G+D training step (SRCycleGAN.optimize_parameters), batch %d per GPU, %dx%d->%dx%d RGB patches"
                     % (B, lr, lr, lr * 4, lr * 4), metric=METRIC, unit=UNIT, units=B, step_resident=step_resident,
                step_e2e=step_e2e, h2d=host_B.numel() * 4, gflop=GFLOP_PER_PATCH * (lr / 64.0) ** 2, cpu=cpu,
                nets=[model.netG_A, model.netG_B, model.netD_A, model.netD_B], opts=[model.optimizer_G, model.optimizer_D])


def _cascade(args, dev, rank, lab: bool):
    """configs[2] / configs[3]: the cascaded SR + colourisation trainers (src/trainCas.py:133-153; trainCasConstLAB.py)."""
    import torch
    from srcgan_b200 import color, trainer_cas
    opt = trainer_cas.params()
    opt.device = dev
    if lab:     # runConstLAB.sh: SRDN at full resolution (x2 blur), ResDeconv(1, 2) on LAB tensors
        opt.up, opt.SRModel, opt.CModel, opt.variant = 2, "SRDN", "ResDeconv", "ConstLAB"
    else:       # run.sh: RDDBNet x4 + ResDeconv
        opt.up, opt.SRModel, opt.CModel, opt.variant = 4, "RDDBNet", "ResDeconv", ""
    torch.manual_seed(0)
    model = trainer_cas.CasSRC(opt)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    host_A = torch.rand(B, 1, 256, 256, generator=g).pin_memory()
    if lab:     # the dataset hands out uint8 RGB tiles; Basic._arr2lab (dataset.py:148-159) runs on the device here
        host_B = (torch.rand(B, 256, 256, 3, generator=g) * 255).to(torch.uint8).pin_memory()
        to_dev_B = lambda: color.image_to_lab(host_B.to(dev, non_blocking=True))
        h2d = host_A.numel() * 4 + host_B.numel()
    else:
        host_B = torch.rand(B, 3, 256, 256, generator=g).pin_memory()
        to_dev_B = lambda: host_B.to(dev, non_blocking=True)
        h2d = host_A.numel() * 4 + host_B.numel() * 4
    real_A, real_B = host_A.to(dev), to_dev_B()

    def step_resident():
        model.optimize_parameters(real_A, real_B)

    def step_e2e():
        model.init_log()
        model.optimize_parameters(host_A.to(dev, non_blocking=True), to_dev_B())
        return 4 * len(model.log_values())                            # one batched read of the four logged scalars

    cpu = lambda batch: cpu_cascade(lab, batch)

    what = ("SRDN x2 (full-resolution input) + ResDeconv(1,2) on LAB tensors, trainCasConstLAB.py step, uint8 RGB -> LAB on the device"
            if lab else "RDDBNet x4 + ResDeconv, trainCas.py step")
    return dict(name="cascaded SR + colourisation: %s, batch %d per GPU, 256x256 targets" % (what, B),
                metric="cascade train-step patches/sec (%s)" % ("ConstLAB" if lab else "x4"), unit=UNIT, units=B,
                step_resident=step_resident, step_e2e=step_e2e, h2d=h2d, gflop=None, cpu=cpu,
                nets=[model.netG_A2C, model.netG_C2B], opts=[model.optimizer_G, model.optimizer_D])


def wl_eval(args, dev, rank):
    """configs[4]: eval sweep (src/test.py / testCas.py:63-85): generator inference on 512x512 tiles + MSE / PSNR / AE / SSIM."""
    import torch
    import torch.nn.functional as F
    from srcgan_b200 import evaluate, nn as snn
    torch.manual_seed(0)
    net = snn.RDDBNetB(3, 3, 64, nb=3, mode="x4").to(dev).eval()
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    host_hr = [torch.rand(1, 3, 512, 512, generator=g).pin_memory() for _ in range(B)]
    hr = [t.to(dev) for t in host_hr]
    lr = [F.interpolate(t, scale_factor=0.25, mode="nearest") for t in hr]

    lr_all, hr_all = torch.cat(lr), torch.cat(hr)
    tb = args.tile_batch

    def step_resident():
        with torch.no_grad():
            for i in range(0, B, tb):
                evaluate.per_image_metrics(net(lr_all[i:i + tb]), hr_all[i:i + tb])

    def step_e2e():
        pairs = []
        for t in host_hr:
            h = t.to(dev, non_blocking=True)
            pairs.append((F.interpolate(h, scale_factor=0.25, mode="nearest"), h))
        rows, _mean = evaluate.evaluate(net, pairs, batch=tb)          # ONE device -> host copy for the whole sweep
        return 16 * len(rows)

    cpu = lambda batch: cpu_eval()

    return dict(name="eval sweep: RDDBNetB x4 inference on 128x128 -> 512x512 tiles + fused MSE/PSNR/AE/SSIM, %d tiles per step per GPU, "
                     "%d tiles per forward (every tile scored on its own)" % (B, tb),
                metric="eval sweep tiles/sec (512x512, x4)", unit="tiles/s", units=B, step_resident=step_resident,
                step_e2e=step_e2e, h2d=B * 3 * 512 * 512 * 4, gflop=62.904 * 4.0, cpu=cpu, nets=[], opts=[])


WORKLOADS = {"gd": wl_gd, "cascade": lambda a, d, r: _cascade(a, d, r, False), "cascade_lab": lambda a, d, r: _cascade(a, d, r, True),
             "eval": wl_eval}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from srcgan_b200 import _lib, dist as sdist, nn as snn, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the srcgan_b200 arm has no CPU fallback")
    local_rank = sdist.init_from_env()
    rank, world = sdist.rank(), sdist.world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    snn.set_precision(args.precision)
    peaks = load_peaks()
    if args.batch is None:
        args.batch = {"gd": 64, "cascade": 64, "cascade_lab": 8, "eval": 16}[args.workload]
    wl = WORKLOADS[args.workload](args, dev, rank)
    reducer = None
    if world > 1 and wl["nets"]:
        sdist.broadcast_module_state(wl["nets"])
        reducer = sdist.attach(wl["opts"], wl["nets"], overlap=args.overlap, bucketed=not args.no_overlap)
    B = wl["units"]

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

    step_resident = wl["step_resident"]
    d2h_bytes = 0

    def step_e2e():
        nonlocal d2h_bytes
        d2h_bytes = wl["step_e2e"]()

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
    ops.timer.by_shape = args.by_shape
    ops.timer.reset()
    ms_instr = timed(step_resident, args.steps)
    ops.timer.enabled = False
    ksum = ops.timer.summary()
    ops.timer.reset()
    e2e = None
    if not args.no_e2e:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        e2e = {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": wl["unit"],
               "h2d_bytes_per_step": wl["h2d"], "d2h_bytes_per_step": d2h_bytes}
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
        traffic, traffic_note = measured_traffic(name)
        roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops"], "traffic": traffic, "traffic_note": traffic_note,
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
    cfg = {"workload": wl["name"], "global_batch": B * world, "parallelism": "dp%d" % world,
           "l2": "not flushed: each step streams >10 GB of activations per GPU through the 126 MB L2"}
    if wl["gflop"] is not None:
        cfg["algorithmic_gflop_per_unit"] = wl["gflop"]
        cfg["step_tflops"] = value * wl["gflop"] / 1e3 / world
    if reducer is not None and hasattr(reducer, "launched"):
        cfg["allreduce"] = {"overlapped_launches": reducer.launched, "pre_hook_launches": reducer.deferred,
                            "fallbacks": reducer.fallbacks,
                            "how": "one in-place NCCL all_reduce per network over its flat gradient bucket, " +
                                   ("launched inside backward when the network's last wgrad has been issued, joined in the "
                                    "optimizer step pre-hook" if reducer.overlap else
                                    "launched and joined (stream-side) in the optimizer step pre-hook")}
    line = {
        "metric": wl["metric"], "value": value, "unit": wl["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic", "config": cfg,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, dt, kind, sample = wl["cpu"](1)
        line["cpu_baseline"] = {"value": v, "unit": wl["unit"], "cores": cores, "kind": kind, "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


# per-launch DRAM traffic of the dominant kernel family, from the committed `ncu --set full` captures (profiles/): the
# family's launches span several layer shapes, so the record names the shape it was measured on
TRAFFIC = {
    "conv3x3_wgrad_stack_tc": (2.73e9 + 0.02e9, "160->32 @64x256x256 (profiles/r1_ncu_wgrad_stack.txt): dram read 2.73 GB + write 0.02 GB "
                               "per launch vs 1.61 GB algorithmic"),
    "conv3x3_sweep2_tc<32,2>": (0.563e9 + 0.248e9, "64->32 @64x256x256 (profiles/r1_ncu_sweep2.txt): dram read 563 MB + write 248 MB per "
                                "launch vs 805 MB algorithmic"),
    "conv3x3_sweep2_tc<64,2>": (0.897e9 + 0.247e9, "192->64 @32x256x256 (profiles/r1_ncu_sweep2.txt): dram read 897 MB + write 247 MB per "
                                "launch vs 1074 MB algorithmic"),
}


def measured_traffic(kernel: str):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if kernel in t:
            return float(t[kernel]["bytes_per_launch"]), t[kernel]["note"]
    except Exception:
        pass
    if kernel in TRAFFIC:
        return TRAFFIC[kernel]
    return None, "no ncu --set full capture of this kernel family is committed"


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-gpu":
        run_torch_gpu(args)
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
