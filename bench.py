#!/usr/bin/env python
"""Headline benchmark: InT training throughput (fwd + BPTT bwd, gradient all-reduce for N > 1) on
synthetic 64-frame 32x32 PathTracker, 14 distractors, batch 64 per GPU (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (N>1 under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]       # CPU arm: the reference algorithm
                                                                   # (oracle port) on the host cores

One JSON line on stdout (rank 0).  A "step" = one forward + backward pass over one batch of 64 clips
per GPU; the optimizer step is excluded.  `value` = clips/s with inputs resident in HBM;
`e2e` = the same through the nn.Module API from pinned HOST buffers (H2D inside the timed region,
loss read back to the host every step).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train seq/s (fwd+bwd) InT 64-frame PathTracker"
UNIT = "seq/s"
B_PER_GPU, T, HW, DIST, SPEED = 64, 64, 32, 14, 1
C, K = 32, 7
# algorithmic FLOPs (1 MAC = 2 FLOP; recompute not counted)
FWD_FLOP_PER_SEQ = T * HW * HW * (2 * (2 * C * C * K * K) + 6 * (2 * C * C) + 2 * 3 * C) + HW * HW * (2 * C + 100)
BWD_FLOP_PER_SEQ = 2 * T * HW * HW * (2 * (2 * C * C * K * K) + 6 * (2 * C * C)) + T * HW * HW * 6 * C + 2 * HW * HW * (2 * C + 100)
FLOP_PER_SEQ = FWD_FLOP_PER_SEQ + BWD_FLOP_PER_SEQ          # 41.902 GFLOP
CONV_FLOP_PER_LAUNCH = 2 * B_PER_GPU * HW * HW * C * C * K * K  # one 7x7 conv over the per-GPU batch: 6.58 GFLOP
FP32_FFMA_PEAK_TFLOPS = 74.4   # nominal 148 SM x 128 lanes x 2 x 1.965 GHz (not in MEASURED_PEAKS.json)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at the headline shape, from the
# committed `ncu --set full` captures (profiles/r1e_tc_kernels_full.txt, profiles/r1a_conv7_ffma_full.txt)
NCU_TRAFFIC_BYTES = {"bf16x3": 8.650e6, "fp32": 8.623e6}   # dram read + write per conv launch, profiles/r1i_kernels_full.txt


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


def cpu_reference_seq_per_s(steps, warmup, batch=8):
    """the reference algorithm on the host cores: oracle port (torch CPU ops + autograd), fp32,
    fwd + BCE + bwd on a bounded sample of the headline workload (batch `batch` instead of 64)."""
    import torch
    import torch.nn.functional as F
    from oracle import int_oracle as O
    from oracle.synth_oracle import synth
    from pathtracker_models_b200 import InT
    torch.set_num_threads(_host_cores())   # torchrun exports OMP_NUM_THREADS=1: the CPU arm still gets every core
    torch.manual_seed(0)
    m = InT.InT(dimensions=C, timesteps=T, kernel_size=K)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    p = O.unpack(sd, dtype=torch.float32)
    for v in p.values():
        v.requires_grad_(True)
    xn, yn = synth(1234, 0, batch, T, HW, HW, DIST, SPEED)
    x, y = torch.from_numpy(xn), torch.from_numpy(yn).view(batch, 1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        logits, _ = O.forward(x, p, keep=False)
        F.binary_cross_entropy_with_logits(logits, y).backward()
        for v in p.values():
            v.grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return batch * len(times) / tot, tot / len(times), torch.get_num_threads(), \
        f"InT fwd+bwd, batch {batch} (of 64), T={T}, {HW}x{HW}, fp32, {len(times)} timed passes after {warmup} warm-up"


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
    print(json.dumps({"impl": "reference-gpu-eager", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "steps": steps,
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
    v, sec, cores, sample = cpu_reference_seq_per_s(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "InT 64 frames 32x32, 14 distractors, batch 64/GPU (BASELINE configs[1])",
                       "note": "CPU arm: each step is a bounded sample (batch 8) of that workload on the host cores"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from pathtracker_models_b200 import InT, _lib, ffhgru_hierarchy, synth_pathtracker
    from pathtracker_models_b200.parallel import attach_data_parallel, broadcast_parameters, shard_first_index

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
    B, T, HW = args.batch, args.frames, args.size      # defaults = the headline workload
    flop_per_seq = (T * HW * HW * (2 * (2 * C * C * K * K) + 6 * (2 * C * C) + 2 * 3 * C) + HW * HW * (2 * C + 100)
                    + 2 * T * HW * HW * (2 * (2 * C * C * K * K) + 6 * (2 * C * C)) + T * HW * HW * 6 * C
                    + 2 * HW * HW * (2 * C + 100))

    torch.manual_seed(0)
    if args.model == "convlstm":
        return run_convlstm(args, dev, world, rank)
    if args.model == "gru":
        return run_gru(args, dev, world, rank)
    if args.eval:
        return run_eval(args, dev, world, rank)
    ctor = InT.InT if args.model == "int" else ffhgru_hierarchy.FFhGRU
    model = ctor(dimensions=C, timesteps=T, kernel_size=K).to(dev)
    model.conv_mode = {"bf16x3": _lib.PTK_CONV_BF16X3, "fp32": _lib.PTK_CONV_FP32, "bf16": _lib.PTK_CONV_BF16}[args.conv_mode]
    if world > 1:
        broadcast_parameters(model)
        attach_data_parallel(model)

    # resident inputs: a few different batches (per-step working set -- 2.2 GB of checkpoints -- exceeds L2)
    nbuf = 4
    batches = [synth_pathtracker(B, T, HW, HW, n_distractors=DIST, speed=SPEED, seed=1234,
                                 first_index=shard_first_index(i, rank, world, B), device=dev) for i in range(nbuf)]

    def step(x, y):
        model.zero_grad(set_to_none=True)
        out, _ = model(x)
        loss = F.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(*batches[i % nbuf])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()

    def timed_region(profile):
        launches0 = _lib.launch_count()
        if profile:
            _lib.profile_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(*batches[i % nbuf])
        e1.record()
        barrier()
        prof = _lib.profile_stop() if profile else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), _lib.launch_count() - launches0, prof

    # pass 1: the headline number, nothing but the K steps between the two events
    ms_total, launches, _ = timed_region(profile=False)
    value = world * B * steps / (ms_total / 1e3)
    # pass 2: the same K steps with every kernel-class launch bracketed by CUDA events on its stream
    # (event records between kernels cost a few percent, so this pass only feeds the roofline block)
    ms_prof, _, prof = timed_region(profile=True)

    # ---- end to end: pinned host float32 clips (channels_last_3d strides, as utils/engine.py:222-243) -> H2D
    # inside the timed region, loss read back each step
    hx = [b[0].permute(0, 2, 3, 4, 1).contiguous().cpu().pin_memory() for b in batches]
    hy = [b[1].cpu().pin_memory() for b in batches]

    from pathtracker_models_b200.data import DevicePrefetcher

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_run(n):
        """every step: its clip goes host -> device (one batch in flight on a copy stream) and its loss comes back to
        the host.  The loss is copied asynchronously into pinned memory and READ one step later (how a training loop
        logs it), so the device never idles on the host; the last value is read after the loop."""
        seen = []
        for i, (x, y) in enumerate(DevicePrefetcher(((hx[j % nbuf], hy[j % nbuf]) for j in range(n)), dev)):
            loss = step(x.permute(0, 4, 1, 2, 3), y)
            loss_host[i & 1:(i & 1) + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                seen.append(float(loss_host[(i - 1) & 1]))
        loss_ev[(n - 1) & 1].synchronize()
        seen.append(float(loss_host[(n - 1) & 1]))
        assert len(seen) == n and all(v == v for v in seen)

    e2e_run(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(steps)
    f1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * steps / (ms2.item() / 1e3)

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
        conv_ms, conv_n = prof["conv_fwd"][0] + prof["conv_dgrad"][0], prof["conv_fwd"][1] + prof["conv_dgrad"][1]
        conv_flop = 2 * B * HW * HW * C * C * K * K
        conv_avg_ms = conv_ms / max(conv_n, 1)
        achieved = conv_flop / (conv_avg_ms * 1e-3) / 1e12 if conv_n else None
        shares = {k: round(v[0] / ms_prof, 4) for k, v in prof.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.conv_mode == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "%s %d frames %dx%d, 14 distractors, batch %d/GPU%s" % (
                           "InT" if args.model == "int" else "FFhGRU", T, HW, HW, B,
                           " (BASELINE configs[1])" if (args.model, T, HW, B) == ("int", 64, 32, 64) else ""),
                       "global_batch": B * world, "parallelism": "dp%d" % world, "conv_mode": args.conv_mode,
                       "l2": "no flush: per-step working set (GBs of checkpoints per GPU) exceeds the 126 MB L2; "
                             "%d distinct input batches rotate" % nbuf},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(hx[0].numel() * 4 + hy[0].numel() * 4),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "optimizer": optimizer_block,
            "clocks": sampler.result(),
            "roofline": {"bound": "tensor", "kernel": ("conv7_ffma_kernel" if args.conv_mode == "fp32" else "conv7_tc_kernel") + " (fwd + dgrad launches)",
                         "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": (achieved / tensor_peak) if achieved else None,
                         "traffic": NCU_TRAFFIC_BYTES.get(args.conv_mode) if (args.model, T, HW, B) == ("int", 64, 32, 64) else None,
                         "traffic_unit": "bytes/launch (ncu --set full, profiles/r1i_kernels_full.txt); algorithmic: 16.8 MB "
                                         "(8.4 MB in + 8.4 MB out, the output stays in L2)",
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % src,
                         "launch_ms": conv_avg_ms, "launches_timed": conv_n, "flop_per_launch": conv_flop,
                         "frac_of_fp32_ffma_nominal": (achieved / FP32_FFMA_PEAK_TFLOPS) if achieved else None,
                         "note": ("3-pass split-bf16 tcgen05 conv: issues 3x the algorithmic MACs, so the fp32-parity ceiling "
                                  "is 1/3 of the bf16 peak (frac_of_third_peak)" if args.conv_mode == "bf16x3" else
                                  "strict-fp32 CUDA-core conv: its own pipe peaks at 74.4 TFLOP/s nominal"),
                         "frac_of_third_peak": (3 * achieved / tensor_peak) if achieved else None,
                         "step_tflops": value / world * flop_per_seq / 1e12,
                         "step_frac_of_tensor_peak": value / world * flop_per_seq / 1e12 / tensor_peak,
                         "kernel_time_share": shares,
                         "measured": "second pass of the same %d steps with per-kernel CUDA events (%.2f ms/step "
                                     "instrumented vs %.2f ms/step plain)" % (steps, ms_prof / steps, ms_total / steps)},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, sec, cores, sample = cpu_reference_seq_per_s(steps=16, warmup=1)     # ~10-15 s of host work
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_convlstm(args, dev, world, rank):
    """BASELINE configs[4]: ConvLSTM(timesteps=64, filt_size=15), x [64,1,32,32], CrossEntropy, train mode
    (so the Jacobian-penalty branch runs, as in the reference).  Secondary line, not the headline."""
    import numpy as np
    import torch
    from pathtracker_models_b200 import _lib, convlstm, synth_pathtracker
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    B, T, HW = args.batch, args.frames, args.size
    rs = np.random.RandomState(0)
    gabor = torch.from_numpy(rs.normal(0, 0.3, (25, 1, 7, 7)).astype(np.float32))   # the .npy is not on the GPU box
    torch.manual_seed(0)
    m = convlstm.ConvLSTM(timesteps=T, filt_size=15, gabor=gabor).to(dev).train()
    m.conv_mode = _lib.PTK_CONV_FP32 if args.conv_mode == "fp32" else _lib.PTK_CONV_BF16X3
    clips, y = synth_pathtracker(B, 2, HW, HW, device=dev)
    x = clips[:, 0:1, 0].contiguous()                                               # [B,1,H,W] static image
    tgt = (clips[:, 2, 0] > 0).long()
    crit = torch.nn.CrossEntropyLoss()

    def step():
        m.zero_grad(set_to_none=True)
        out, pen, loss = m(x, 0, 0, tgt, crit)
        loss.backward()
        return loss

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
    ms = e0.elapsed_time(e1)
    flop = 3 * 2.0 * B * HW * HW * 100 * 25 * 225 * T          # fwd + dgrad + wgrad of the recurrent conv
    print(json.dumps({"metric": "train seq/s (fwd+bwd) ConvLSTM", "value": B * steps / (ms / 1e3), "unit": UNIT,
                      "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "ConvLSTM timesteps=%d filt 15, x [%d,1,%d,%d] (BASELINE configs[4])" % (T, B, HW, HW)},
                      "gpu_launches": _lib.launch_count() - l0,
                      "roofline": {"bound": "tensor", "achieved": flop * steps / (ms / 1e3) / 1e12, "unit": "TFLOP/s",
                                   "peak": load_peaks()[0], "frac": flop * steps / (ms / 1e3) / 1e12 / load_peaks()[0],
                                   "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s); the 3-pass split issues 3x "
                                                  "the algorithmic MACs, so the fp32-parity ceiling is 1/3 of it" % load_peaks()[2],
                                   "note": "recurrent-conv FLOPs only; conv_mode %s (bf16x3: forward / transposed convolution and "
                                           "weight gradient on tcgen05, 3-pass split-bf16)" % ("fp32" if args.conv_mode == "fp32" else "bf16x3")}}),
          flush=True)


def run_gru(args, dev, world, rank):
    """kys.GRU (`--model gru`, utils/engine.py:147-153: 64 channels, 7x7) on the headline clip shape.  Secondary line:
    first version of this family, strict-fp32 CUDA-core convolutions."""
    import torch
    import torch.nn.functional as F
    from pathtracker_models_b200 import _lib, kys, synth_pathtracker
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    B, T, HW = args.batch, args.frames, args.size
    torch.manual_seed(0)
    m = kys.GRU(dimensions=64, timesteps=T, kernel_size=7).to(dev)
    m.conv_mode = {"bf16x3": _lib.PTK_CONV_BF16X3, "fp32": _lib.PTK_CONV_FP32, "bf16": _lib.PTK_CONV_BF16}[args.conv_mode]
    x, y = synth_pathtracker(B, T, HW, HW, n_distractors=DIST, speed=SPEED, device=dev)

    def step():
        m.zero_grad(set_to_none=True)
        out, _ = m(x)
        loss = F.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        return loss

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
    ms = e0.elapsed_time(e1)
    flop = 3 * 2.0 * B * HW * HW * 49 * (128 * 128 + 128 * 64) * T     # fwd + dgrad + wgrad of the two cell convolutions
    print(json.dumps({"metric": "train seq/s (fwd+bwd) kys.GRU", "value": B * steps / (ms / 1e3), "unit": UNIT,
                      "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "kys.GRU %d frames %dx%d, batch %d" % (T, HW, HW, B)},
                      "gpu_launches": _lib.launch_count() - l0,
                      "roofline": {"bound": "tensor", "achieved": flop * steps / (ms / 1e3) / 1e12, "unit": "TFLOP/s",
                                   "peak": load_peaks()[0], "frac": flop * steps / (ms / 1e3) / 1e12 / load_peaks()[0],
                                   "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s); fp32-parity ceiling = 1/3 of it "
                                                  "(3-pass split-bf16)" % load_peaks()[2],
                                   "note": "cell-convolution FLOPs only; conv_mode %s (bf16x3: wide tcgen05 kernels shared with the "
                                           "ConvLSTM)" % args.conv_mode}}), flush=True)


def run_eval(args, dev, world, rank):
    """forward-only (inference) throughput: model.eval(), no_grad -- BASELINE configs[2] asks for it at 128 frames 64x64."""
    import torch
    from pathtracker_models_b200 import InT, _lib, ffhgru_hierarchy, synth_pathtracker
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    B, T, HW = args.batch, args.frames, args.size
    torch.manual_seed(0)
    ctor = InT.InT if args.model == "int" else ffhgru_hierarchy.FFhGRU
    m = ctor(dimensions=C, timesteps=T, kernel_size=K).to(dev).eval()
    m.conv_mode = {"bf16x3": _lib.PTK_CONV_BF16X3, "fp32": _lib.PTK_CONV_FP32, "bf16": _lib.PTK_CONV_BF16}[args.conv_mode]
    x, _ = synth_pathtracker(B, T, HW, HW, n_distractors=DIST, speed=SPEED, device=dev)
    with torch.no_grad():
        for _ in range(warmup):
            m(x)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"metric": "eval seq/s (fwd only)", "value": B * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                      "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "bf16" if args.conv_mode == "bf16" else "f32", "data": "synthetic",
                      "config": {"workload": "%s %d frames %dx%d, batch %d, forward only" % (args.model, T, HW, HW, B),
                                 "conv_mode": args.conv_mode}, "gpu_launches": _lib.launch_count() - l0}), flush=True)


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
    ap.add_argument("--model", default="int", choices=["int", "hgru", "convlstm", "gru"], help="non-headline configs (BASELINE configs[2..3])")
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
