"""bench.py -- MoCo pretrain images/sec on N B200s (BASELINE.json metric), plus the kernel roofline.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm: the oracle port of the reference step on host cores

A step = one MoCo iteration (train.py:244-283): query encoder fwd, ShuffleBN permute, key encoder fwd,
un-shuffle, q.Queue^T + InfoNCE + dq, enqueue, backward, SGD step, EMA update -- ResNet-50, feat_dim 128,
batch 256/GPU, bf16 autocast, synthetic 224x224 images, random-init weights.
Workloads: N=1 -> BASELINE configs[1] (K=16384); N>1 -> configs[2] (K=65536, ShuffleBN over NVLink P2P).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--feat-dim", type=int, default=128)
    ap.add_argument("--nce-k", type=int, default=0, help="queue length (0: 16384 at 1 GPU, 65536 otherwise)")
    ap.add_argument("--nce-t", type=float, default=0.07)
    ap.add_argument("--memory-format", default="channels_last", choices=["channels_last", "contiguous"],
                    help="encoder activation layout (host PyTorch side)")
    ap.add_argument("--no-stress", action="store_true", help="skip the c5 roofline-stress microbench")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=8)
    return ap.parse_args()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of nce_stats_kernel from the committed
# `ncu --set full` captures (profiles/r1_final_nce_c2_ncu_metrics.csv, profiles/r1_final_nce_c5_ncu_metrics.csv)
NCU_TRAFFIC_BYTES = {(256, 128, 16384): 4287232, (512, 256, 262144): 134522880 + 3844352,   # statistics kernel
                     # one-pass kernel (profiles/r1_onepass_c2_ncu_metrics.csv, r1_onepass_c5_ncu_metrics.csv)
                     ("onepass", 256, 128, 16384): 4308992, ("onepass", 512, 256, 262144): 134518528 + 3464192}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as exc:       # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_arm(args, steps, warmup):
    from oracle.cpu_step import time_cpu_arm
    K = args.nce_k or 16384
    r = time_cpu_arm(args.arch, args.feat_dim, K, args.nce_t, args.cpu_sample_batch, steps, warmup)
    return r, K


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    r, K = cpu_arm(args, args.steps, args.warmup)
    cores = r["threads"]
    line = {
        "impl": "reference", "metric": "MoCo pretrain images/sec", "value": r["images_per_s"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.arch} feat_dim={args.feat_dim} K={K} batch=256/GPU (BASELINE configs[1]); "
                               f"CPU arm runs a bounded sample of {r['batch']} images/step"},
        "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{r['steps']} steps x {r['batch']} images, oracle/cpu_step.py (numpy hot path + "
                                   f"torch-CPU fp32 encoders), {cores} threads"},
        "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def stress_roofline(peaks, dev):
    """BASELINE configs[4]: N=512, C=256, K=262144 -- the tensor-bound shape, hot-path kernels alone (queue 134 MB > L2).
    Default path = ONE sweep over the queue producing loss statistics AND dq (4NCK FLOP); the two-pass alternative
    (statistics kernel, then dq kernel that recomputes S: 6NCK FLOP executed, 4NCK credited) is timed beside it."""
    import torch
    import torch.nn.functional as F
    from moco_b200 import _lib
    lib = _lib.load()
    N, C, K, T = 512, 256, 262144, 0.07
    g = torch.Generator(device=dev).manual_seed(3)
    q = F.normalize(torch.randn(N, C, device=dev, generator=g), dim=1).bfloat16()
    k = F.normalize(torch.randn(N, C, device=dev, generator=g), dim=1).bfloat16()
    queue = F.normalize(torch.randn(K, C, device=dev, generator=g), dim=1).bfloat16()
    f32 = dict(dtype=torch.float32, device=dev)
    lse, lr, pr, lp, dq = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(2, **f32), torch.zeros(N, C, **f32)
    wsb = lib.moco_nce_workspace_bytes(N, C, K)
    ws = torch.zeros(wsb + 256, dtype=torch.uint8, device=dev)
    wp = ws.data_ptr() + (-ws.data_ptr()) % 256
    stream = torch.cuda.current_stream().cuda_stream
    iters = 20

    def timed_kernels(flags):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(iters)]
        for e4 in ev:       # events must exist (be recorded once) before the library records into them; stop before
            for j in (1, 0, 3, 2):   # start, so a hook that never fires reads as a NEGATIVE interval
                e4[j].record()

        def call(i):
            if i >= 0:
                lib.moco_prof_set_events(1, ev[i][0].cuda_event, ev[i][1].cuda_event)
                lib.moco_prof_set_events(2, ev[i][2].cuda_event, ev[i][3].cuda_event)
            rc = lib.moco_nce_fwd(q.data_ptr(), k.data_ptr(), 1, queue.data_ptr(), N, C, K, 1.0 / T, None, lse.data_ptr(),
                                  lr.data_ptr(), pr.data_ptr(), lp.data_ptr(), dq.data_ptr(), wp, wsb, flags, stream)
            _lib.check(rc, "moco_nce_fwd")
        for _ in range(3):
            call(-1)
        for i in range(iters):
            call(i)
        lib.moco_prof_set_events(1, None, None)
        lib.moco_prof_set_events(2, None, None)
        torch.cuda.synchronize()
        return (sum(e[0].elapsed_time(e[1]) for e in ev) * 1e3 / iters, sum(e[2].elapsed_time(e[3]) for e in ev) * 1e3 / iters)

    _, us_one = timed_kernels(_lib.NCE_AUTO)                      # one-pass kernel reports on the DQ hook
    us_stats, us_dq = timed_kernels(_lib.NCE_TWO_PASS)
    flops = 2.0 * N * C * (K + 1)                                 # per direction: fwd = bwd = 2NC(K+1)
    bytes_ = K * C * 2 + 3 * N * C * 2 + 12 * N
    a = 2 * flops / (us_one * 1e-6) / 1e12
    return {
        "workload": "BASELINE configs[4]: N=512 feat_dim=256 K=262144 (hot-path kernels alone, queue 134 MB > L2)",
        "kernel": "nce_dq2_kernel<FUSED> (one sweep: S=q.Queue^T, P=2^(S/T-m), O+=P.Queue, row sums) -> loss + dq",
        "bound": "tensor", "achieved": a, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": a / peaks["tf_burst"],
        "us_per_launch": us_one, "algorithmic_flops": 2 * flops, "hbm_GBps": bytes_ / (us_one * 1e-6) / 1e9,
        "traffic": NCU_TRAFFIC_BYTES.get(("onepass", N, C, K)),
        "two_pass": {"stats_kernel_us": us_stats, "stats_TFLOPs": flops / (us_stats * 1e-6) / 1e12,
                     "stats_frac": flops / (us_stats * 1e-6) / 1e12 / peaks["tf_burst"],
                     "dq_kernel_us": us_dq, "dq_TFLOPs_executed": 2 * flops / (us_dq * 1e-6) / 1e12,
                     "sum_us": us_stats + us_dq,
                     "note": "statistics pass + dq pass (recomputes S): 6NCK executed for the same 4NCK of algorithmic work"},
    }


def run_native(args):
    import torch
    import torch.distributed as dist
    from moco_b200 import _lib, encoders
    from moco_b200.NCE import MemoryMoCo
    from moco_b200.train_step import MoCoStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = load_peaks()
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True

    N, C, T = args.batch, args.feat_dim, args.nce_t
    K = args.nce_k or (16384 if world == 1 else 65536)
    torch.manual_seed(0)
    ctor = getattr(encoders, args.arch)
    mf = torch.channels_last if args.memory_format == "channels_last" else torch.contiguous_format
    model = ctor(low_dim=C).to(dev).to(memory_format=mf)
    model_ema = ctor(low_dim=C).to(dev).to(memory_format=mf)
    model_ema.load_state_dict(model.state_dict())
    contrast = MemoryMoCo(C, K, T).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.03 * N * world / 256, momentum=0.9, weight_decay=1e-4)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], broadcast_buffers=False)
    nhwc = args.memory_format == "channels_last"
    step = MoCoStep(model, model_ema, contrast, opt, channels_last=nhwc)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    dev_inputs = torch.randn(N, 6, 224, 224, device=dev, generator=gen)           # dataset.py:31-33 layout
    host_inputs = [torch.empty(N, 6, 224, 224, pin_memory=True).copy_(dev_inputs) for _ in range(2)]
    epoch = 1

    def split(t):
        x1, x2 = torch.split(t, [3, 3], dim=1)                       # train.py:250 (views of the 6-channel batch)
        if nhwc:
            return x1, x2                 # MoCoStep reads the crops in place (moco_crop_to_nhwc_bf16)
        return (x1.contiguous(memory_format=mf), x2.contiguous())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- arm 1: inputs resident in HBM (308 MB per step > 126 MB L2)
    x1, x2 = split(dev_inputs)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for e4 in ev:
        for j in (1, 0, 3, 2):       # stop before start: a hook that never fires reads as a negative interval
            e4[j].record()

    def loop_resident(steps, profile=False):
        for i in range(steps):
            if profile:
                lib.moco_prof_set_events(1, ev[i][0].cuda_event, ev[i][1].cuda_event)
                lib.moco_prof_set_events(2, ev[i][2].cuda_event, ev[i][3].cuda_event)
            step(x1, x2, epoch)

    loop_resident(args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launches
    ms_total = timed(lambda s: loop_resident(s, True), args.steps)
    launches = _lib.launches - l0
    sampler.stop_flag = True
    sampler.join()
    lib.moco_prof_set_events(1, None, None)
    lib.moco_prof_set_events(2, None, None)
    us_stats = sum(e[0].elapsed_time(e[1]) for e in ev) * 1e3 / args.steps
    us_dq = sum(e[2].elapsed_time(e[3]) for e in ev) * 1e3 / args.steps
    ms_step = ms_total / args.steps
    value = N * world * args.steps / (ms_total * 1e-3)

    # ---- arm 2: end to end through the public API with HOST inputs (pinned), H2D inside the timed region,
    #      next batch prefetched on a copy stream, loss + prob read back every step
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(dev_inputs) for _ in range(2)]
    sink = []

    def loop_e2e(steps):
        main = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            bufs[0].copy_(host_inputs[0], non_blocking=True)
        for i in range(steps):
            main.wait_stream(copy_stream)
            cur = bufs[i & 1]
            if i + 1 < steps:
                copy_stream.wait_stream(main)         # buffer (i+1)&1 was consumed by step i-1
                with torch.cuda.stream(copy_stream):
                    bufs[(i + 1) & 1].copy_(host_inputs[(i + 1) & 1], non_blocking=True)
            a, b = split(cur)
            loss, prob = step(a, b, epoch)
            sink.append((loss.item(), prob.item()))    # D2H read of the step's result (train.py:280-281)

    loop_e2e(max(2, args.warmup // 2))
    ms_e2e = timed(loop_e2e, args.steps)
    e2e_value = N * world * args.steps / (ms_e2e * 1e-3)
    h2d = N * 6 * 224 * 224 * 4
    final_loss = sink[-1][0]

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # dominant hot-path kernel inside the step: the one-pass kernel (T = 0.07 -> MOCO_NCE_AUTO takes one sweep);
    # algorithmic work per launch = forward 2NC(K+1) + backward 2NC(K+1) FLOP, queue read once
    flops = 4.0 * N * C * (K + 1)
    bytes_ = K * C * 2 + 3 * N * C * 2 + 12 * N
    one_pass = us_stats <= 0.0                        # the statistics-kernel hook never fired
    us_main = us_dq if one_pass else us_stats + us_dq
    a_tf = flops / (us_main * 1e-6) / 1e12
    roofline = {
        "kernel": ("nce_dq2_kernel<FUSED>: one sweep over the queue on tcgen05 (S = q.Queue^T, P = 2^(S/T - m), "
                   "O += P.Queue, row sums) -> loss statistics + dq" if one_pass else
                   "nce_stats_kernel + nce_dq2_kernel (two-pass)") + ", timed inside the step",
        "bound": "tensor", "achieved": a_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
        "frac": a_tf / peaks["tf_sustained"], "peak_source": peaks["source"] + ", sustained bf16",
        "us_per_launch": us_main, "algorithmic_flops": flops, "algorithmic_bytes": bytes_,
        "hbm_GBps": bytes_ / (us_main * 1e-6) / 1e9, "hbm_frac": bytes_ / (us_main * 1e-6) / 1e9 / peaks["hbm_gbs"],
        "traffic": NCU_TRAFFIC_BYTES.get(("onepass", N, C, K)),
        "note": f"ideal time for this shape is {flops / (peaks['tf_sustained'] * 1e12) * 1e6:.1f} us "
                f"({flops / 1e9:.2f} GFLOP, {bytes_ / 1e6:.1f} MB): two 128-row tiles per CTA, so launch + prologue + "
                "one pipeline fill dominate; roofline_stress (N=1 runs) is the tensor-bound shape of BASELINE configs[4]",
    }
    line = {
        "metric": "MoCo pretrain images/sec (device-timed, max over ranks)", "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.arch} feat_dim={C} K={K} batch={N}/GPU bf16 "
                               + ("(BASELINE configs[1])" if world == 1 else "(BASELINE configs[2], ShuffleBN P2P permute)"),
                   "global_batch": N * world, "parallelism": f"dp{world}", "temperature": T,
                   "l2": "inputs (308 MB/step) exceed L2; no explicit flush"},
        "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
        "gpu_launches": launches,
        "roofline": roofline,
        "final_loss": final_loss,
    }
    if world == 1 and not args.no_stress:
        line["roofline_stress"] = stress_roofline(peaks, dev)
    if world == 1 and not args.no_cpu_baseline:
        r, Kc = cpu_arm(args, 2, 1)
        line["cpu_baseline"] = {"value": r["images_per_s"], "unit": "images/s", "cores": r["threads"], "kind": "port",
                                "sample": f"{r['steps']} steps x {r['batch']} images (same {args.arch}, K={Kc}), "
                                          f"oracle/cpu_step.py on {r['threads']} host threads"}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
