#!/usr/bin/env python
"""Headline benchmark: InT training throughput (fwd + BPTT bwd, gradient all-reduce for N > 1) on
synthetic 64-frame 32x32 PathTracker, 14 distractors, batch 64 per GPU (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (N>1 under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]       # CPU arm: the reference algorithm
                                                                   # (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" = one forward + backward pass over one batch of 64 clips
per GPU; the optimizer step is excluded.  `value` = clips/s with inputs resident in HBM;
`e2e` = the same through the nn.Module API from pinned HOST buffers (H2D inside the timed region,
loss read back to the host every step).  At N = 1 the line also carries
  `parity`         one un-timed check of loss + every gradient of the CUDA path against the CPU port on the
                   cpu_baseline sample (same weights, same clips),
  `other_configs`  BASELINE.json configs[2]-[4] (InT 128 frames 64x64; FFhGRU fp32-parity and bf16; ConvLSTM),
                   each with value / ms_per_step / e2e / roofline / cpu_baseline, short runs (--no-extra skips them).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "seq/s"
B_PER_GPU, T, HW, DIST, SPEED = 64, 64, 32, 14, 1
C, K = 32, 7
FP32_FFMA_PEAK_TFLOPS = 74.4   # nominal 148 SM x 128 lanes x 2 x 1.965 GHz (not in MEASURED_PEAKS.json)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at the headline shape, from the
# committed `ncu --set full` capture named in NCU_TRAFFIC_SOURCE (bench.py cannot run ncu on itself)
NCU_TRAFFIC_BYTES = {"bf16x3": 8.665e6, "fp32": 8.623e6}
NCU_TRAFFIC_SOURCE = "profiles/r5z_kernels_full.txt (conv7_tc2_kernel forward launch, final build: 8.67 MB; fp32: profiles/r1a_conv7_ffma_full.txt)"


def int_flop_per_seq(T_, HW_):
    """algorithmic FLOPs of one InT / FFhGRU sequence, fwd + bwd (1 MAC = 2 FLOP; recompute not counted)"""
    px = HW_ * HW_
    fwd = T_ * px * (2 * (2 * C * C * K * K) + 6 * (2 * C * C) + 2 * 3 * C) + px * (2 * C + 100)
    bwd = 2 * T_ * px * (2 * (2 * C * C * K * K) + 6 * (2 * C * C)) + T_ * px * 6 * C + 2 * px * (2 * C + 100)
    return fwd + bwd


def metric_name(model, frames, eval_only=False):
    fam = {"int": "InT", "hgru": "FFhGRU", "convlstm": "ConvLSTM", "gru": "kys.GRU"}[model]
    if eval_only:
        return "eval seq/s (fwd only) %s %d-frame PathTracker" % (fam, frames)
    return "train seq/s (fwd+bwd) %s %d-frame PathTracker" % (fam, frames)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                     "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("sampler_error:" + type(e).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _host_cores():
    """physical cores this process may run on (what torch picks by itself when OMP_NUM_THREADS is unset)."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or avail
    except Exception:  # noqa: BLE001
        phys = avail
    return max(1, min(avail, phys))


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference ALGORITHM on the host cores.  /root/reference is Python and absent on the GPU box, so this is
# the oracle port (oracle/int_oracle.py: torch CPU ops + autograd, fp32; 1x1 gates as einsum), `kind: "port"`.
def cpu_port_int(steps, warmup, batch=8, frames=T, size=HW, variant="int", keep_last=False, fp64=False):
    """fwd + BCE + bwd of the InT / FFhGRU restatement on a bounded sample of the workload (batch `batch`).
    returns (seq/s, s/pass, threads, sample description[, loss, grads by state_dict key])"""
    import torch
    import torch.nn.functional as F
    from oracle import int_oracle as O
    from oracle.synth_oracle import synth
    from pathtracker_models_b200 import InT, ffhgru_hierarchy
    torch.set_num_threads(_host_cores())   # torchrun exports OMP_NUM_THREADS=1: the CPU arm still gets every core
    torch.manual_seed(0)
    ctor = InT.InT if variant == "int" else ffhgru_hierarchy.FFhGRU
    m = ctor(dimensions=C, timesteps=frames, kernel_size=K)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    dt_ = torch.float64 if fp64 else torch.float32
    p = O.unpack(sd, dtype=dt_)
    for v in p.values():
        v.requires_grad_(True)
    xn, yn = synth(1234, 0, batch, frames, size, size, DIST, SPEED)
    x, y = torch.from_numpy(xn).to(dt_), torch.from_numpy(yn).view(batch, 1).to(dt_)
    times, last = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        logits, _ = O.forward(x, p, variant, keep=False)
        loss = F.binary_cross_entropy_with_logits(logits, y)
        loss.backward()
        if keep_last and i == warmup + steps - 1:
            last = (loss.item(), O.grad_to_state_keys({k: v.grad.clone() for k, v in p.items() if v.grad is not None}, sd))
        for v in p.values():
            v.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    desc = "%s fwd+bwd, batch %d, T=%d, %dx%d, fp32, %d timed passes after %d warm-up" % (
        "InT" if variant == "int" else "FFhGRU", batch, frames, size, size, len(times), warmup)
    res = (batch * len(times) / tot, tot / len(times), torch.get_num_threads(), desc)
    return res + last if keep_last else res


def cpu_port_convlstm(steps, warmup, batch, frames, size, gabor):
    import torch
    import torch.nn.functional as F
    from oracle import convlstm_oracle as CO
    from pathtracker_models_b200 import convlstm
    torch.set_num_threads(_host_cores())
    torch.manual_seed(0)
    m = convlstm.ConvLSTM(timesteps=frames, filt_size=15, gabor=gabor)
    p = CO.unpack(m.state_dict(), dtype=torch.float32)
    for v in p.values():
        v.requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(batch, 1, size, size, generator=g)
    tgt = (torch.rand(batch, size, size, generator=g) > 0.5).long()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out, _ = CO.forward(x, p, frames)
        F.cross_entropy(out, tgt).backward()
        for v in p.values():
            v.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return batch * len(times) / tot, tot / len(times), torch.get_num_threads(), \
        "ConvLSTM filt 15 fwd+bwd (input-path convolutions hoisted, no penalty branch), batch %d, timesteps=%d, %dx%d, fp32, " \
        "%d timed passes after %d warm-up" % (batch, frames, size, size, len(times), warmup)


def gpu_eager_reference(args):
    """SURVEY 8(d) "honest beat-this bar": the reference's OPERATOR SEQUENCE (oracle/int_eager.py: F.conv2d, F.batch_norm,
    one ATen kernel per point-wise operator, autograd tape) run by PyTorch eager on cuda:0 at the full headline shape,
    TF32 off (fp32 parity class) and on.  Extra information only: `--impl reference` without --ref-device stays the CPU arm."""
    import torch
    import torch.nn.functional as F
    from oracle import int_eager
    from oracle import int_oracle as O
    from pathtracker_models_b200 import InT, synth_pathtracker
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = InT.InT(dimensions=C, timesteps=T, kernel_size=K)
    p = {k: v.to(dev).requires_grad_(True) for k, v in O.unpack(m.state_dict(), dtype=torch.float32).items()}
    batches = [synth_pathtracker(args.batch, T, HW, HW, n_distractors=DIST, first_index=i * args.batch, device=dev)
               for i in range(2)]
    out = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                e0.record()
            x, y = batches[i % 2]
            F.binary_cross_entropy_with_logits(int_eager.forward(x, p), y).backward()
            for v in p.values():
                v.grad = None
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out["tf32_on" if tf32 else "tf32_off"] = {"value": args.batch / (ms / 1e3), "ms_per_step": ms}
    print(json.dumps({"impl": "reference-gpu-eager", "metric": metric_name("int", T), "unit": UNIT, "n_gpus": 1, "steps": steps,
                      "warmup": warmup, "value": out["tf32_off"]["value"], "ms_per_step": out["tf32_off"]["ms_per_step"],
                      "tf32_on": out["tf32_on"], "tf32_off": out["tf32_off"], "dtype": "f32", "data": "synthetic",
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                      "config": {"workload": "InT 64 frames 32x32, 14 distractors, batch %d/GPU (BASELINE configs[1])" % args.batch,
                                 "note": "oracle/int_eager.py = the reference's operator sequence in PyTorch eager on the same B200; "
                                         "not the product path, not the driver's reference arm"}}), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    v, sec, cores, sample = cpu_port_int(steps, warmup)
    line = {"impl": "reference", "metric": metric_name("int", T), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "InT 64 frames 32x32, 14 distractors, batch 64/GPU (BASELINE configs[1])",
                       "note": "CPU arm: the reference ALGORITHM as restated in oracle/int_oracle.py (torch CPU ops + autograd; the "
                               "reference itself is Python and is not on the GPU box); each step is a bounded sample (batch 8 of "
                               "64) of that workload on the host cores"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
class Timer:
    """device timing of K steps on the current stream, max over ranks"""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(self, fn, n):
        import torch
        import torch.distributed as dist
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        fn(n)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()


def make_e2e_runner(step, host_batches, dev, to_device_view=None):
    """every step: its inputs go pinned host -> device (one batch in flight on a copy stream) and its loss comes back to
    the host.  The loss is copied asynchronously into pinned memory and READ one step later (how a training loop logs
    it), so the device never idles on the host; the last value is read after the loop."""
    import torch
    from pathtracker_models_b200.data import DevicePrefetcher
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    nb = len(host_batches)

    def e2e_run(n):
        seen = []
        for i, batch in enumerate(DevicePrefetcher((host_batches[j % nb] for j in range(n)), dev)):
            if to_device_view is not None:
                batch = to_device_view(batch)
            loss = step(*batch)
            loss_host[i & 1:(i & 1) + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                seen.append(float(loss_host[(i - 1) & 1]))
        loss_ev[(n - 1) & 1].synchronize()
        seen.append(float(loss_host[(n - 1) & 1]))
        assert len(seen) == n and all(v == v for v in seen)

    return e2e_run


def measure(step, batches, host_batches, steps, warmup, timer, per_step_units, to_device_view=None):
    """resident-input throughput + end-to-end throughput of `step` (same K steps each); returns a dict"""
    from pathtracker_models_b200 import _lib
    nb = len(batches)
    for i in range(warmup):
        step(*batches[i % nb])
    l0 = _lib.launch_count()
    ms = timer.run(lambda n: [step(*batches[i % nb]) for i in range(n)], steps)
    launches = _lib.launch_count() - l0
    e2e_run = make_e2e_runner(step, host_batches, timer.dev, to_device_view)
    e2e_run(2)
    ms2 = timer.run(e2e_run, steps)
    h2d = int(sum(t.numel() * t.element_size() for t in host_batches[0]))
    return {"value": timer.world * per_step_units * steps / (ms / 1e3), "ms_per_step": ms / steps, "gpu_launches": launches,
            "e2e": {"value": timer.world * per_step_units * steps / (ms2 / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4}}


def int_batches(B, T_, HW_, rank, world, dev, nbuf):
    from pathtracker_models_b200 import synth_pathtracker
    from pathtracker_models_b200.parallel import shard_first_index
    batches = [synth_pathtracker(B, T_, HW_, HW_, n_distractors=DIST, speed=SPEED, seed=1234,
                                 first_index=shard_first_index(i, rank, world, B), device=dev) for i in range(nbuf)]
    # host clips: pinned float32 with channels_last_3d strides, as utils/engine.py:222-243 hands them over
    host = [(b[0].permute(0, 2, 3, 4, 1).contiguous().cpu().pin_memory(), b[1].cpu().pin_memory()) for b in batches]
    return batches, host


def parity_vs_cpu_port(model, dev):
    """one un-timed pass of the CUDA path on the cpu_baseline sample (batch 8, same seed-0 weights, same clips) against
    the CPU port: loss and every parameter gradient.  Also times the port (this IS the cpu_baseline leg)."""
    import torch
    import torch.nn.functional as F
    from oracle.synth_oracle import synth
    v, sec, cores, sample, loss32, g32 = cpu_port_int(steps=12, warmup=1, keep_last=True)
    _, _, _, _, loss_cpu, g_cpu = cpu_port_int(steps=1, warmup=0, keep_last=True, fp64=True)      # the arbiter
    xn, yn = synth(1234, 0, 8, T, HW, HW, DIST, SPEED)
    x, y = torch.from_numpy(xn).to(dev), torch.from_numpy(yn).view(8, 1).to(dev)
    model.zero_grad(set_to_none=True)
    out, _ = model(x)
    loss = F.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    rel = lambda a, b: (a.double().cpu() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-300)
    worst, worst_name, worst_ratio, ratio_name = 0.0, None, 0.0, None
    for n, p in model.named_parameters():
        if p.grad is None or n not in g_cpu or float(g_cpu[n].abs().max()) == 0.0:
            continue
        e, noise = rel(p.grad, g_cpu[n]), rel(g32[n], g_cpu[n])
        if e > worst:
            worst, worst_name = e, n
        r = e / max(1e-4, 4 * noise)
        if r > worst_ratio:
            worst_ratio, ratio_name = r, n
    model.zero_grad(set_to_none=True)
    parity = {"sample": "batch 8, T=64, 32x32 (the cpu_baseline sample), seed-0 default-init weights",
              "checker": "CPU port in fp64 (oracle/int_oracle.py); the fp32 port run gives the fp32 noise floor per tensor",
              "loss_abs_err": abs(loss.item() - loss_cpu), "loss_abs_err_fp32_port": abs(loss32 - loss_cpu),
              "grad_rel_err_max": worst, "grad_worst_tensor": worst_name,
              "worst_fraction_of_allowed": worst_ratio, "tightest_tensor": ratio_name, "pass": bool(worst_ratio <= 1.0),
              "rule": "per tensor max|a-b|/max|b| vs fp64 <= max(1e-4, 4 x the fp32 port's own error) (SURVEY 8c); the same gate "
                      "runs in tests/ on T=64 / T=128 fixtures generated by the reference itself"}
    return parity, {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}


def run_other_configs(args, dev, timer, tensor_peak, peak_src):
    """BASELINE.json configs[2]-[4] at N = 1: short runs, same measurement rules as the headline."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from pathtracker_models_b200 import InT, _lib, convlstm, ffhgru_hierarchy
    steps, warmup = 4, 3
    out = []

    def roof(value, flop_per_unit, third):
        tf = value * flop_per_unit / 1e12
        r = {"bound": "tensor", "achieved": tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tf / tensor_peak,
             "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % peak_src, "traffic": None,
             "note": "whole step, algorithmic FLOPs / step time"}
        if third:
            r["frac_of_third_peak"] = 3 * tf / tensor_peak
            r["note"] += "; 3-pass split-bf16: fp32-parity ceiling = 1/3 of the bf16 peak"
        return r

    def int_like(tag, ctor, B, T_, HW_, mode, baseline_cfg, cpu_batch):
        torch.manual_seed(0)
        model = ctor(dimensions=C, timesteps=T_, kernel_size=K).to(dev)
        model.conv_mode = {"bf16x3": _lib.PTK_CONV_BF16X3, "bf16": _lib.PTK_CONV_BF16}[mode]
        batches, host = int_batches(B, T_, HW_, 0, 1, dev, 2)

        def step(x, y):
            model.zero_grad(set_to_none=True)
            o, _ = model(x)
            loss = F.binary_cross_entropy_with_logits(o, y)
            loss.backward()
            return loss

        r = measure(step, batches, host, steps, warmup, timer, B, lambda b: (b[0].permute(0, 4, 1, 2, 3), b[1]))
        fam = "int" if ctor is InT.InT else "hgru"
        r.update({"metric": metric_name(fam, T_), "unit": UNIT, "steps": steps, "warmup": warmup,
                  "dtype": "bf16" if mode == "bf16" else "f32",
                  "config": {"workload": "%s %d frames %dx%d, 14 distractors, batch %d/GPU (%s)" % (
                      "InT" if fam == "int" else "FFhGRU", T_, HW_, HW_, B, baseline_cfg), "conv_mode": mode},
                  "roofline": roof(r["value"], int_flop_per_seq(T_, HW_), mode == "bf16x3")})
        if cpu_batch and not args.no_cpu_baseline:
            v, sec, cores, sample = cpu_port_int(steps=2, warmup=1, batch=cpu_batch, frames=T_, size=HW_, variant=fam)
            r["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        del model, batches, host
        torch.cuda.empty_cache()
        return r

    out.append(int_like("cfg3", InT.InT, 16, 128, 64, "bf16x3", "BASELINE configs[2]", 1))
    out.append(int_like("cfg4", ffhgru_hierarchy.FFhGRU, B_PER_GPU, T, HW, "bf16x3", "BASELINE configs[3], fp32-parity variant", 4))
    out.append(int_like("cfg4b", ffhgru_hierarchy.FFhGRU, B_PER_GPU, T, HW, "bf16", "BASELINE configs[3], bf16 variant: single-pass "
                        "bf16 MMAs, not fp32-parity", 0))
    out.append(int_like("cfg2b", InT.InT, B_PER_GPU, T, HW, "bf16", "BASELINE configs[1] shape, bf16 variant: single-pass bf16 MMAs, "
                        "not fp32-parity", 0))

    # ConvLSTM (BASELINE configs[4]): timesteps = 64, filt 15, x [64,1,32,32], CrossEntropy, train mode (penalty branch runs)
    gpath = os.path.join(ROOT, "tests", "golden", "convlstm_k7_small.npz")
    gabor = torch.from_numpy(np.load(gpath)["gabor"]).float()       # the reference's utils/gabor_serre.npy, kept as a fixture
    torch.manual_seed(0)
    m = convlstm.ConvLSTM(timesteps=T, filt_size=15, gabor=gabor).to(dev).train()
    m.conv_mode = _lib.PTK_CONV_BF16X3
    g = torch.Generator().manual_seed(1)
    host = [(torch.rand(B_PER_GPU, 1, HW, HW, generator=g).pin_memory(),
             (torch.rand(B_PER_GPU, HW, HW, generator=g) > 0.5).long().pin_memory()) for _ in range(2)]
    batches = [(h[0].to(dev), h[1].to(dev)) for h in host]
    crit = torch.nn.CrossEntropyLoss()

    def lstm_step(x, tgt):
        m.zero_grad(set_to_none=True)
        o, pen, loss = m(x, 0, 0, tgt, crit)
        loss.backward()
        return loss

    r = measure(lstm_step, batches, host, steps, warmup, timer, B_PER_GPU)
    flop = 3 * 2.0 * HW * HW * 100 * 25 * 225 * T          # fwd + dgrad + wgrad of the recurrent convolution, per sequence
    r.update({"metric": metric_name("convlstm", T), "unit": UNIT, "steps": steps, "warmup": warmup, "dtype": "f32",
              "config": {"workload": "ConvLSTM timesteps=%d filt 15, x [%d,1,%d,%d], train mode incl. the Jacobian-penalty branch "
                                     "(BASELINE configs[4])" % (T, B_PER_GPU, HW, HW), "conv_mode": "bf16x3",
                         "gabor": "reference bank (tests/golden fixture)"},
              "roofline": roof(r["value"], flop, True)})
    r["roofline"]["note"] += "; recurrent-convolution FLOPs only"
    if not args.no_cpu_baseline:
        v, sec, cores, sample = cpu_port_convlstm(steps=1, warmup=0, batch=1, frames=T, size=HW, gabor=gabor)
        r["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    out.append(r)
    # ... and its bf16 variant: the same tcgen05 kernels issuing one bf16 product per k-step (not fp32-parity)
    m.conv_mode = _lib.PTK_CONV_BF16
    r2 = measure(lstm_step, batches, host, steps, warmup, timer, B_PER_GPU)
    r2.update({"metric": metric_name("convlstm", T), "unit": UNIT, "steps": steps, "warmup": warmup, "dtype": "f32",
               "config": {"workload": "ConvLSTM timesteps=%d filt 15, x [%d,1,%d,%d], train mode (BASELINE configs[4] shape, bf16 variant: "
                                      "single-pass bf16 MMAs in the recurrent convolutions, not fp32-parity)" % (T, B_PER_GPU, HW, HW),
                          "conv_mode": "bf16", "gabor": "reference bank (tests/golden fixture)"},
               "roofline": roof(r2["value"], flop, False)})
    r2["roofline"]["note"] += "; recurrent-convolution FLOPs only"
    out.append(r2)
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from pathtracker_models_b200 import InT, _lib, ffhgru_hierarchy
    from pathtracker_models_b200.parallel import attach_data_parallel, broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    B, T_, HW_ = args.batch, args.frames, args.size      # defaults = the headline workload
    flop_per_seq = int_flop_per_seq(T_, HW_)
    timer = Timer(dev, world)

    torch.manual_seed(0)
    if args.model == "convlstm":
        return run_convlstm(args, dev, world, rank)
    if args.model == "gru":
        return run_gru(args, dev, world, rank)
    if args.eval:
        return run_eval(args, dev, world, rank)
    ctor = InT.InT if args.model == "int" else ffhgru_hierarchy.FFhGRU
    model = ctor(dimensions=C, timesteps=T_, kernel_size=K).to(dev)
    model.conv_mode = {"bf16x3": _lib.PTK_CONV_BF16X3, "fp32": _lib.PTK_CONV_FP32, "bf16": _lib.PTK_CONV_BF16}[args.conv_mode]
    if world > 1:
        broadcast_parameters(model)
        attach_data_parallel(model)

    # resident inputs: a few different batches (per-step working set -- GBs of checkpoints -- exceeds L2)
    nbuf = 4
    batches, host = int_batches(B, T_, HW_, rank, world, dev, nbuf)

    def step(x, y):
        model.zero_grad(set_to_none=True)
        out, _ = model(x)
        loss = F.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        return loss

    for i in range(warmup):
        step(*batches[i % nbuf])
    timer.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # pass 1 (+ the end-to-end pass): the headline numbers, nothing but the K steps between the two events
    main = measure(step, batches, host, steps, 0, timer, B, lambda b: (b[0].permute(0, 4, 1, 2, 3), b[1]))
    # pass 2: the same K steps with every kernel-class launch bracketed by CUDA events on its stream
    # (event records between kernels cost a few percent, so this pass only feeds the roofline block)
    _lib.profile_start()
    ms_prof = timer.run(lambda n: [step(*batches[i % nbuf]) for i in range(n)], steps)
    prof = _lib.profile_stop()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    grad_check = None
    if world > 1:
        # once per run: every rank must hold the same averaged gradient after the all-reduce (cross-rank checksum)
        step(*batches[0])
        flat = torch.cat([p.grad.reshape(-1).double() for p in model.parameters() if p.grad is not None])
        cs = torch.stack([flat.sum(), flat.abs().sum()])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        grad_check = {"checksum_sum": cs[0].item(), "checksum_abs": cs[1].item(),
                      "max_rank_spread": (hi - lo).abs().max().item(), "identical_on_all_ranks": bool(torch.equal(lo, hi))}

    headline = (args.model, T_, HW_, B) == ("int", 64, 32, 64)
    parity = cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and headline:
        # BEFORE the optimizer timing below touches the weights: the CPU port rebuilds the same seed-0 model
        parity, cpu_base = parity_vs_cpu_port(model, dev)     # ~15 s of host work

    # ---- the step either side of the path, reported separately ("optimizer step excluded and reported
    # separately"): loss + metrics kernel and fused Adam + zero_grad vs torch.optim.Adam on the same gradients
    from pathtracker_models_b200.optim import FusedAdam, bce_with_logits_and_scores

    def time_ms(fn, n=10):
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / n

    step(*batches[0])
    saved_grads = [None if p.grad is None else p.grad.clone() for p in model.parameters()]

    def restore():
        for p, g in zip(model.parameters(), saved_grads):
            p.grad = None if g is None else g.clone()

    def opt_time(opt, fused):
        def one():
            restore()
            if fused:
                opt.step(zero_grad=True)
            else:
                opt.step()
                opt.zero_grad()
        one()
        return time_ms(one) - time_ms(restore)

    torch_adam_ms = opt_time(torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-6), False)
    fused_adam_ms = opt_time(FusedAdam(model.parameters(), lr=1e-6), True)
    zl, yl = torch.randn(B, 1, device=dev), batches[0][1]
    loss_ms = time_ms(lambda: bce_with_logits_and_scores(zl, yl))
    optimizer_block = {"fused_adam_zero_grad_ms": round(fused_adam_ms, 4), "torch_adam_zero_grad_ms": round(torch_adam_ms, 4),
                       "bce_scores_ms": round(loss_ms, 4), "note": "outside the timed region; lr 1e-6 so the benchmark weights barely move"}

    if rank == 0:
        tensor_peak, hbm_peak, src = load_peaks()
        # the dominant kernel = the 7x7 convolution.  Its FORWARD launches are pure convolution (+ BatchNorm partial statistics);
        # the data-gradient launches of the tensor-core modes also carry the fused BatchNorm backward (a second operand stream and
        # the dY write-back), so they are reported beside it (dgrad_launch_ms) instead of being averaged into the roofline figure
        conv_ms, conv_n = prof["conv_fwd"][0], prof["conv_fwd"][1]
        conv_flop = 2 * B * HW_ * HW_ * C * C * K * K
        conv_avg_ms = conv_ms / max(conv_n, 1)
        dgrad_avg_ms = prof["conv_dgrad"][0] / max(prof["conv_dgrad"][1], 1)
        achieved = conv_flop / (conv_avg_ms * 1e-3) / 1e12 if conv_n else None
        shares = {k: round(v[0] / ms_prof, 4) for k, v in prof.items()}
        value = main["value"]
        line = {
            "metric": metric_name(args.model, T_), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.conv_mode == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "%s %d frames %dx%d, 14 distractors, batch %d/GPU%s" % (
                           "InT" if args.model == "int" else "FFhGRU", T_, HW_, HW_, B, " (BASELINE configs[1])" if headline else ""),
                       "global_batch": B * world, "parallelism": "dp%d" % world, "conv_mode": args.conv_mode,
                       "l2": "no flush: per-step working set (GBs of checkpoints per GPU) exceeds the 126 MB L2; "
                             "%d distinct input batches rotate" % nbuf},
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "optimizer": optimizer_block,
            "clocks": sampler.result(),
            "roofline": {"bound": "tensor", "kernel": ("conv7_ffma_kernel" if args.conv_mode == "fp32" else "conv7_tc2_kernel") + " (forward launches)",
                         "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": (achieved / tensor_peak) if achieved else None,
                         "traffic": NCU_TRAFFIC_BYTES.get(args.conv_mode) if headline else None,
                         "traffic_unit": "bytes/launch (dram read + write, ncu --set full, %s); algorithmic: 16.8 MB "
                                         "(8.4 MB in + 8.4 MB out, the output stays in L2)" % NCU_TRAFFIC_SOURCE,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % src,
                         "launch_ms": conv_avg_ms, "launches_timed": conv_n, "flop_per_launch": conv_flop,
                         "dgrad_launch_ms": dgrad_avg_ms,
                         "frac_of_fp32_ffma_nominal": (achieved / FP32_FFMA_PEAK_TFLOPS) if achieved else None,
                         "note": ("3-pass split-bf16 tcgen05 conv: issues 3x the algorithmic MACs, so the fp32-parity ceiling "
                                  "is 1/3 of the bf16 peak (frac_of_third_peak)" if args.conv_mode == "bf16x3" else
                                  "strict-fp32 CUDA-core conv: its own pipe peaks at 74.4 TFLOP/s nominal" if args.conv_mode == "fp32"
                                  else "single-pass bf16 conv"),
                         "frac_of_third_peak": (3 * achieved / tensor_peak) if achieved else None,
                         "step_tflops": value / world * flop_per_seq / 1e12,
                         "step_frac_of_tensor_peak": value / world * flop_per_seq / 1e12 / tensor_peak,
                         "kernel_time_share": shares,
                         "measured": "second pass of the same %d steps with per-kernel CUDA events (%.2f ms/step "
                                     "instrumented vs %.2f ms/step plain)" % (steps, ms_prof / steps, main["ms_per_step"])},
        }
        if grad_check is not None:
            line["multi_gpu_gradient_check"] = grad_check
        if parity is not None:
            line["parity"], line["cpu_baseline"] = parity, cpu_base
        elif world == 1 and not args.no_cpu_baseline:
            v, sec, cores, sample = cpu_port_int(steps=4, warmup=1, batch=max(1, min(8, 2048 // (HW_ * HW_ // 16))),
                                                 frames=T_, size=HW_, variant=args.model)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        if world == 1 and headline and not args.no_extra:
            del batches, host
            torch.cuda.empty_cache()
            line["other_configs"] = run_other_configs(args, dev, timer, tensor_peak, src)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _simple_run(step, steps, warmup):
    import torch
    from pathtracker_models_b200 import _lib
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), _lib.launch_count() - l0


def run_convlstm(args, dev, world, rank):
    """BASELINE configs[4]: ConvLSTM(timesteps=64, filt_size=15), x [64,1,32,32], CrossEntropy, train mode
    (so the Jacobian-penalty branch runs, as in the reference).  Secondary line, not the headline."""
    import numpy as np
    import torch
    from pathtracker_models_b200 import _lib, convlstm, synth_pathtracker
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    B, T_, HW_ = args.batch, args.frames, args.size
    gabor = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "convlstm_k7_small.npz"))["gabor"]).float()
    torch.manual_seed(0)
    jp = bool(getattr(args, "jacobian_penalty", False))      # the penalty inside the graph: loss + 10 * penalty.mean() (mainclean.py:192-196)
    m = convlstm.ConvLSTM(timesteps=T_, filt_size=15, gabor=gabor, jacobian_penalty=jp).to(dev).train()
    m.conv_mode = {"fp32": _lib.PTK_CONV_FP32, "bf16": _lib.PTK_CONV_BF16}.get(args.conv_mode, _lib.PTK_CONV_BF16X3)
    clips, y = synth_pathtracker(B, 2, HW_, HW_, device=dev)
    x = clips[:, 0:1, 0].contiguous()                                               # [B,1,H,W] static image
    tgt = (clips[:, 2, 0] > 0).long()
    crit = torch.nn.CrossEntropyLoss()

    def step():
        m.zero_grad(set_to_none=True)
        out, pen, loss = m(x, 0, 0, tgt, crit)
        if jp:
            loss = loss + pen.mean() * 1e1
        loss.backward()
        return loss

    ms, launches = _simple_run(step, steps, warmup)
    flop = 3 * 2.0 * B * HW_ * HW_ * 100 * 25 * 225 * T_          # fwd + dgrad + wgrad of the recurrent conv
    peak, _, src = load_peaks()
    print(json.dumps({"metric": metric_name("convlstm", T_), "value": B * steps / (ms / 1e3), "unit": UNIT,
                      "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "ConvLSTM timesteps=%d filt 15, x [%d,1,%d,%d] (BASELINE configs[4])%s"
                                             % (T_, B, HW_, HW_, ", jacobian_penalty=True: loss + 10 * penalty.mean()" if jp else ""),
                                 "conv_mode": args.conv_mode if args.conv_mode in ("fp32", "bf16") else "bf16x3"},
                      "gpu_launches": launches,
                      "roofline": {"bound": "tensor", "achieved": flop * steps / (ms / 1e3) / 1e12, "unit": "TFLOP/s",
                                   "peak": peak, "frac": flop * steps / (ms / 1e3) / 1e12 / peak,
                                   "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s); the 3-pass split issues 3x "
                                                  "the algorithmic MACs, so the fp32-parity ceiling is 1/3 of it" % src,
                                   "note": "recurrent-conv FLOPs only"}}), flush=True)


def run_gru(args, dev, world, rank):
    """kys.GRU (`--model gru`, utils/engine.py:147-153: 64 channels, 7x7) on the headline clip shape.  Secondary line."""
    import torch
    import torch.nn.functional as F
    from pathtracker_models_b200 import _lib, kys, synth_pathtracker
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    B, T_, HW_ = args.batch, args.frames, args.size
    torch.manual_seed(0)
    m = kys.GRU(dimensions=64, timesteps=T_, kernel_size=7).to(dev)
    m.conv_mode = {"bf16x3": _lib.PTK_CONV_BF16X3, "fp32": _lib.PTK_CONV_FP32, "bf16": _lib.PTK_CONV_BF16}[args.conv_mode]
    x, y = synth_pathtracker(B, T_, HW_, HW_, n_distractors=DIST, speed=SPEED, device=dev)

    def step():
        m.zero_grad(set_to_none=True)
        out, _ = m(x)
        loss = F.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        return loss

    ms, launches = _simple_run(step, steps, warmup)
    flop = 3 * 2.0 * B * HW_ * HW_ * 49 * (128 * 128 + 128 * 64) * T_     # fwd + dgrad + wgrad of the two cell convolutions
    peak, _, src = load_peaks()
    print(json.dumps({"metric": metric_name("gru", T_), "value": B * steps / (ms / 1e3), "unit": UNIT,
                      "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "kys.GRU %d frames %dx%d, batch %d" % (T_, HW_, HW_, B), "conv_mode": args.conv_mode},
                      "gpu_launches": launches,
                      "roofline": {"bound": "tensor", "achieved": flop * steps / (ms / 1e3) / 1e12, "unit": "TFLOP/s",
                                   "peak": peak, "frac": flop * steps / (ms / 1e3) / 1e12 / peak,
                                   "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s); fp32-parity ceiling = 1/3 of it "
                                                  "(3-pass split-bf16)" % src,
                                   "note": "cell-convolution FLOPs only"}}), flush=True)


def run_eval(args, dev, world, rank):
    """forward-only (inference) throughput: model.eval(), no_grad -- BASELINE configs[2] asks for it at 128 frames 64x64."""
    import torch
    from pathtracker_models_b200 import InT, _lib, ffhgru_hierarchy, synth_pathtracker
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    B, T_, HW_ = args.batch, args.frames, args.size
    torch.manual_seed(0)
    ctor = InT.InT if args.model == "int" else ffhgru_hierarchy.FFhGRU
    m = ctor(dimensions=C, timesteps=T_, kernel_size=K).to(dev).eval()
    m.conv_mode = {"bf16x3": _lib.PTK_CONV_BF16X3, "fp32": _lib.PTK_CONV_FP32, "bf16": _lib.PTK_CONV_BF16}[args.conv_mode]
    x, _ = synth_pathtracker(B, T_, HW_, HW_, n_distractors=DIST, speed=SPEED, device=dev)
    with torch.no_grad():
        ms, launches = _simple_run(lambda: m(x), steps, warmup)
    print(json.dumps({"metric": metric_name(args.model, T_, eval_only=True), "value": B * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                      "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "bf16" if args.conv_mode == "bf16" else "f32", "data": "synthetic",
                      "config": {"workload": "%s %d frames %dx%d, batch %d, forward only" % (args.model, T_, HW_, HW_, B),
                                 "conv_mode": args.conv_mode}, "gpu_launches": launches}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="with --impl reference: cuda = additionally time the reference's operator sequence in PyTorch eager on the GPU")
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_configs block (BASELINE configs[2]-[4]) of the default run")
    ap.add_argument("--jacobian-penalty", action="store_true", help="--model convlstm: jacobian_penalty=True, back-propagate loss + 10 * penalty.mean()")
    ap.add_argument("--model", default="int", choices=["int", "hgru", "convlstm", "gru"], help="non-headline configs (BASELINE configs[2..4])")
    ap.add_argument("--eval", action="store_true", help="forward-only throughput (secondary line)")
    ap.add_argument("--frames", type=int, default=T)
    ap.add_argument("--size", type=int, default=HW)
    ap.add_argument("--conv-mode", default="bf16x3", choices=["fp32", "bf16x3", "bf16"],
                    help="7x7 convolutions: strict-fp32 FFMA, or 3-pass split-bf16 tcgen05 (both parity-gated)")
    args = ap.parse_args()
    if args.impl == "reference" and args.ref_device == "cuda":
        gpu_eager_reference(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
