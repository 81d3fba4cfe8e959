"""Launches every support kernel of the hot path once per iteration at the BASELINE shapes, for ncu captures
(`ncu --set full -k regex:...`) and CUDA-event timings of kernels that have no hook of their own.

    python tools/support_kernels.py [iters]          # prints one JSON line of per-kernel CUDA-event times

Kernels: crop_to_nhwc (256 x 3 of 6 x 224 x 224 fp32 -> bf16 NHWC), gather_bulk (256 bf16 image rows, W = 1),
gather_small (2048 x 128 fp32 feature rows), gather_ldg (256 fp32 image rows), crop_gather (the W = 1 ShuffleBN
kernel), enqueue (2048 x 128 into K = 65536), ema_multi (ResNet-50), and the head chain at configs[1] / configs[2] (whatever kernels moco_nce_fwd launches).
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    import torch.nn.functional as F
    from moco_b200 import _lib, encoders
    from moco_b200.NCE import MemoryMoCo
    from moco_b200.util import crop_to_channels_last_bf16, moment_update
    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > L2 (126 MB)

    six = torch.randn(256, 6, 224, 224, device=dev, generator=g)
    img = torch.randn(256, 3, 224, 224, device=dev, generator=g).bfloat16()
    img_out = torch.empty_like(img)
    perm = torch.randperm(256, device=dev)
    feats = torch.randn(2048, 128, device=dev, generator=g)
    feats_out = torch.empty_like(feats)
    perm2 = torch.randperm(2048, device=dev)
    model, ema = encoders.resnet50(low_dim=128).to(dev), encoders.resnet50(low_dim=128).to(dev)
    contrast = {K: MemoryMoCo(128, K, 0.07).to(dev) for K in (16384, 65536)}
    q = F.normalize(torch.randn(256, 128, device=dev, generator=g), dim=1).requires_grad_(True)
    k = F.normalize(torch.randn(256, 128, device=dev, generator=g), dim=1)
    k_all = {16384: k, 65536: F.normalize(torch.randn(2048, 128, device=dev, generator=g), dim=1)}
    k_all[65536][:256] = k
    tab1 = (ctypes.c_void_p * 1)(img.data_ptr())
    tab2 = (ctypes.c_void_p * 1)(feats.data_ptr())
    img32 = six[:, :3].contiguous()
    img32_out = torch.empty_like(img32)
    tab3 = (ctypes.c_void_p * 1)(img32.data_ptr())

    ops = {
        "crop_to_nhwc (256x3x224x224 f32 crop of a 6-channel batch -> bf16 NHWC)": lambda: crop_to_channels_last_bf16(six[:, 3:]),
        "gather_bulk (256 bf16 image rows of 301056 B, W=1)": lambda: _lib.check(lib.moco_shuffle_gather(
            tab1, 1, 256, perm.data_ptr(), 256, 3 * 224 * 224 * 2, img_out.data_ptr(), 0, stream), "gather"),
        "gather_small (2048 f32 feature rows of 512 B)": lambda: _lib.check(lib.moco_shuffle_gather(
            tab2, 1, 2048, perm2.data_ptr(), 2048, 512, feats_out.data_ptr(), 0, stream), "gather"),
        "ema_multi (ResNet-50, 161 tensors, 23.8 M params)": lambda: moment_update(model, ema, 0.999),
        "gather_ldg (256 f32 image rows of 602112 B, W=1, load/store variant)": lambda: _lib.check(lib.moco_shuffle_gather(
            tab3, 1, 256, perm.data_ptr(), 256, 3 * 224 * 224 * 4, img32_out.data_ptr(), 1, stream), "gather"),
        "crop_gather (W=1 ShuffleBN: 256 x 3 of 6 channels f32 -> permuted bf16 NHWC, one kernel)":
            lambda: _lib.check(lib.moco_crop_gather_nhwc_bf16(six[:, 3:].data_ptr(), _lib.dtype_code(six), six.stride(0),
                                                              perm.data_ptr(), img_out.data_ptr(), 256, 3, 224 * 224,
                                                              stream), "crop_gather"),
        "enqueue (2048 x 128 f32 keys into K=65536, bf16 + f32 queues)": lambda: contrast[65536].enqueue(k_all[65536]),
    }

    def head(K):
        def run():
            q.grad = None
            loss, _ = contrast[K].forward_loss(q, k, k_all[K])        # head chain + enqueue
            loss.backward()
        return run
    ops["head chain + enqueue, configs[1] (N=256 C=128 K=16384)"] = head(16384)
    ops["head chain + enqueue, configs[2] (N=256 C=128 K=65536, 2048 keys)"] = head(65536)

    res = {}
    for name, fn in ops.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()                                                  # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = tot * 1e3 / iters
    bytes_ = {"crop": 256 * 3 * 224 * 224 * 6, "gather_bulk": 256 * 3 * 224 * 224 * 2 * 2, "ema": 23770304 * 12}
    res["GBps"] = {"crop_to_nhwc": bytes_["crop"] / res[list(ops)[0]] / 1e3,
                   "gather_bulk": bytes_["gather_bulk"] / res[list(ops)[1]] / 1e3,
                   "ema_multi": bytes_["ema"] / res[list(ops)[3]] / 1e3,
                   "gather_ldg": bytes_["gather_bulk"] * 2 / res[list(ops)[4]] / 1e3,
                   "crop_gather": bytes_["crop"] / res[list(ops)[5]] / 1e3}
    res["note"] = "us per call, CUDA events on the launching stream, L2 flushed before every call (host launch included)"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
