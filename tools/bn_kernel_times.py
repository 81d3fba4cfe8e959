"""Per-kernel GPU time of the four BatchNorm kernels over the ResNet-50 layer shapes (batch 256), from the CUPTI kernel
records of torch.profiler: one line with the per-encoder-pass totals (layer counts applied).

    python tools/bn_kernel_times.py [iters]
"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bn_lab import R50  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    from moco_b200.bn import BatchNormAct2d
    from torch.profiler import profile, ProfilerActivity
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    tot = collections.defaultdict(float)
    per_shape = {}
    for hw, C, relu, has_res, count in R50:
        shape = (256, C, hw, hw)
        cl = torch.channels_last
        x = torch.randn(shape, device=dev, generator=g).bfloat16().contiguous(memory_format=cl).requires_grad_(True)
        res = torch.randn(shape, device=dev, generator=g).bfloat16().contiguous(memory_format=cl).requires_grad_(True) if has_res else None
        dy = torch.randn(shape, device=dev, generator=g).bfloat16().contiguous(memory_format=cl)
        mod = BatchNormAct2d(C, relu=bool(relu)).to(dev)
        for _ in range(2):
            mod(x, res).backward(dy)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(iters):
                mod(x, res).backward(dy)
            torch.cuda.synchronize()
        here = collections.defaultdict(float)
        for ev in prof.events():
            name = ev.name
            for k in ("bn_stats", "bn_apply", "bn_bwd_reduce", "bn_bwd_apply"):
                if k + "_kernel" in name:
                    here[k] += ev.device_time / iters
        per_shape[f"{C}x{hw}{'r' if has_res else ''}"] = {k: round(v, 1) for k, v in here.items()}
        for k, v in here.items():
            tot[k] += v * count
        del x, res, dy
        torch.cuda.empty_cache()
    print(json.dumps({"variant": os.environ.get("MOCO_BN_VARIANT", "0"), "per_encoder_pass_us": {k: round(v, 1) for k, v in tot.items()},
                      "per_shape_us": per_shape}))


if __name__ == "__main__":
    main()
