""""the reference PyTorch path on the same B200 as the primary beat-this baseline for every
kernel".  Times, on one GPU with CUDA events, the post-encoder head (a7-a11: logits + loss + prob + backward to
q + enqueue) and the EMA update (f1) two ways:

  * `torch`  : the reference's sequence of PyTorch library calls restated here op for op
               (Contrast.py:20-34, NCECriterion.py:11-13, train.py:264,273, util.py:124-127) -- fp32 as the
               reference runs it without Apex, and with a bf16 mm (what Apex O1 / autocast would do);
  * `native` : moco_b200 through the C ABI (MemoryMoCo.forward_loss + backward, util.moment_update).

Prints one JSON line per case.  Measurement tool only; nothing in the product imports it.

    python tools/torch_gpu_baseline.py [c2 c3 c5 ema]
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HEAD = {"c1": (32, 128, 1024), "c2": (256, 128, 16384), "c3": (256, 128, 65536), "c4shard": (2048, 128, 16384),
        "c5": (512, 256, 262144)}
T = 0.07


def time_cuda(fn, iters, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)                                    # > L2 (126 MB): evict the queue between iterations
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        total += s.elapsed_time(e)
    return total / iters * 1e3                               # us


class TorchHead:
    """The reference head as PyTorch library calls (op-for-op restatement for timing)."""

    def __init__(self, C, K, mm_dtype):
        self.K, self.index, self.mm_dtype = K, 0, mm_dtype
        stdv = 1.0 / (C / 3) ** 0.5
        self.memory = torch.rand(K, C, device="cuda").mul_(2 * stdv).add_(-stdv)

    def step(self, q, k, k_all):
        k = k.detach()
        l_pos = (q * k).sum(dim=-1, keepdim=True)
        mem = self.memory.clone().detach()
        if self.mm_dtype is torch.float32:
            l_neg = torch.mm(q, mem.transpose(1, 0))
        else:
            l_neg = torch.mm(q.to(self.mm_dtype), mem.to(self.mm_dtype).transpose(1, 0)).float()
        out = torch.cat((l_pos, l_neg), dim=1) / T
        out = out.contiguous()
        with torch.no_grad():
            n_all = k_all.shape[0]
            ids = torch.fmod(torch.arange(n_all, dtype=torch.long) + self.index, self.K).cuda()
            self.memory.index_copy_(0, ids, k_all)
            self.index = (self.index + n_all) % self.K
        label = torch.zeros(out.shape[0], dtype=torch.long, device="cuda")
        loss = F.cross_entropy(out, label)
        prob = F.softmax(out, dim=1)[:, 0].mean()
        loss.backward()
        return loss, prob


def run_head(name):
    from moco_b200.NCE import MemoryMoCo
    N, C, K = HEAD[name]
    torch.manual_seed(0)
    q0 = F.normalize(torch.randn(N, C, device="cuda"), dim=1)
    k = F.normalize(torch.randn(N, C, device="cuda"), dim=1)
    k_all = k.clone()
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda") if K * C * 4 < (200 << 20) else None
    iters = 30 if K <= 65536 else 10
    res = {"case": f"head_{name}", "N": N, "C": C, "K": K, "l2_flush": flush is not None}

    for tag, dt in (("torch_fp32", torch.float32), ("torch_bf16mm", torch.bfloat16)):
        head = TorchHead(C, K, dt)

        def f():
            q = q0.clone().requires_grad_(True)
            head.step(q, k, k_all)
        res[f"{tag}_us"] = round(time_cuda(f, iters, flush=flush), 1)
        del head

    contrast = MemoryMoCo(C, K, T).cuda()

    def g():
        q = q0.clone().requires_grad_(True)
        loss, prob = contrast.forward_loss(q, k, k_all)
        loss.backward()
    res["native_us"] = round(time_cuda(g, iters, flush=flush), 1)
    res["speedup_vs_torch_fp32"] = round(res["torch_fp32_us"] / res["native_us"], 2)
    res["speedup_vs_torch_bf16mm"] = round(res["torch_bf16mm_us"] / res["native_us"], 2)
    print(json.dumps(res), flush=True)


def run_ema(_):
    from moco_b200 import encoders
    from moco_b200.util import moment_update
    model, ema = encoders.resnet50(low_dim=128).cuda(), encoders.resnet50(low_dim=128).cuda()
    n = sum(p.numel() for p in model.parameters())
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")

    def ref():                                               # util.py:124-127 verbatim semantics
        for p1, p2 in zip(model.parameters(), ema.parameters()):
            p2.data.mul_(0.999).add_(p1.detach().data, alpha=1 - 0.999)

    def foreach():
        pe = [p.data for p in ema.parameters()]
        torch._foreach_mul_(pe, 0.999)
        torch._foreach_add_(pe, [p.detach().data for p in model.parameters()], alpha=1 - 0.999)

    res = {"case": "ema_resnet50", "params": n, "tensors": len(list(model.parameters())), "l2_flush": True}
    res["torch_loop_us"] = round(time_cuda(ref, 20, flush=flush), 1)
    res["torch_foreach_us"] = round(time_cuda(foreach, 20, flush=flush), 1)
    res["native_us"] = round(time_cuda(lambda: moment_update(model, ema, 0.999), 20, flush=flush), 1)
    # the kernel alone (the line above includes the Python-side pointer check of 161 parameter pairs)
    from moco_b200 import _lib
    lib, plan = _lib.load(), ema._moco_ema_plan
    st = torch.cuda.current_stream().cuda_stream

    def kern():
        lib.moco_ema_update(plan.segs.data_ptr(), plan.prefix.data_ptr(), plan.n_segs, plan.n_chunks, 0.999, 1 - 0.999, st)
    res["native_kernel_us"] = round(time_cuda(kern, 20, flush=flush), 1)
    res["native_kernel_GBps"] = round(12.0 * n / (res["native_kernel_us"] * 1e-6) / 1e9, 1)
    res["speedup_vs_torch_loop"] = round(res["torch_loop_us"] / res["native_us"], 2)
    print(json.dumps(res), flush=True)


def main():
    assert torch.cuda.is_available()
    names = sys.argv[1:] or ["c2", "c3", "c5", "ema"]
    for n in names:
        (run_ema if n == "ema" else run_head)(n)


if __name__ == "__main__":
    main()
