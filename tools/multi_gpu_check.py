"""Multi-GPU ShuffleBN check + NVLink gather bandwidth (run under torch.distributed.run, one rank per GPU).

Each rank regenerates EVERY rank's batch from per-rank seeds, so it can evaluate the numpy oracle of the
reference's all_gather + index ShuffleBN locally and compare its own P2P-pulled result bit for bit.
Prints one JSON line from rank 0.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from moco_b200 import _lib  # noqa: E402
from moco_b200.util import DistributedShufle, ShuffleContext, dist_collect  # noqa: E402
from oracle import moco_oracle as O  # noqa: E402


def correctness(rank, world, dev):
    """ShuffleBN (both directions, double-buffer reuse, bf16/NHWC publish), dist_collect and three sharded-queue
    steps against the numpy oracle of the reference (util.py:47-111, Contrast.py:20-34).  Also called by bench.py
    before its timed region at N > 1 (the `parity` block of the JSON line).  Collective: every rank calls it."""
    res = {"world": world, "ok": True}

    def batch(r, n, shape, seed):
        g = torch.Generator().manual_seed(seed * 1000 + r)
        return torch.randn(n, *shape, generator=g)

    # ---- correctness: small images + features, several epochs, both directions, repeated (double buffering)
    for it, (n, epoch) in enumerate([(8, 1), (8, 2), (16, 7), (8, 1)]):
        xs = [batch(r, n, (3, 8, 8), 10 + it) for r in range(world)]
        outs, bwd = O.forward_shuffle([x.numpy() for x in xs], epoch)
        mine, binds = DistributedShufle.forward_shuffle(xs[rank].to(dev), epoch)
        ok = np.array_equal(mine.cpu().numpy(), outs[rank]) and np.array_equal(binds.cpu().numpy(), bwd)
        feats = [torch.from_numpy(o.reshape(n, -1)[:, :32].copy()) for o in outs]          # stand-in key encoder
        f_all, f_loc = O.backward_shuffle([f.numpy() for f in feats], bwd, True)
        g_all, g_loc = DistributedShufle.backward_shuffle(feats[rank].to(dev), binds, True)
        ok = ok and np.array_equal(g_all.cpu().numpy(), f_all) and np.array_equal(g_loc.cpu().numpy(), f_loc[rank])
        # S6: un-shuffled local features line up with this rank's own images
        ok = ok and np.array_equal(g_loc.cpu().numpy(), xs[rank].numpy().reshape(n, -1)[:, :32])
        res[f"iter{it}"] = bool(ok)
        res["ok"] = res["ok"] and bool(ok)
    # fused input path: crops of a 6-channel batch published as bf16 NHWC, twice (double buffering)
    for it, epoch in enumerate([3, 4]):
        six = [batch(r, 8, (6, 8, 8), 50 + it) for r in range(world)]
        crops = [O.bf16_round(x[:, 3:].numpy()) for x in six]                              # second crop, bf16 values
        outs, _ = O.forward_shuffle(crops, epoch)
        mine, _ = DistributedShufle.forward_shuffle(six[rank].to(dev)[:, 3:], epoch, channels_last=True)
        ok = (mine.dtype == torch.bfloat16 and mine.is_contiguous(memory_format=torch.channels_last)
              and np.array_equal(mine.float().cpu().numpy(), outs[rank]))
        res[f"nhwc{it}"] = bool(ok)
        res["ok"] = res["ok"] and bool(ok)
        # the same crops as space-to-depth rows (what MoCoStep moves when the encoders take them): the oracle's
        # shuffled bf16 images, re-laid-out with torch ops, must equal the pulled rows bit for bit
        mine, _ = DistributedShufle.forward_shuffle(six[rank].to(dev)[:, 3:], epoch, channels_last="s2d")
        want = torch.from_numpy(outs[rank])                                                  # [n, 3, 8, 8] bf16 values
        xp = torch.nn.functional.pad(want, (4, 2, 4, 2)).view(8, 3, 7, 2, 7, 2).permute(0, 3, 5, 1, 2, 4).reshape(8, 12, 7, 7)
        want = torch.nn.functional.pad(xp, (0, 0, 0, 0, 0, 4))
        ok = (mine.shape == (8, 16, 7, 7) and mine.is_contiguous(memory_format=torch.channels_last)
              and np.array_equal(mine.float().cpu().numpy(), want.numpy()))
        res[f"nhwc_s2d{it}"] = bool(ok)
        res["ok"] = res["ok"] and bool(ok)
    col = dist_collect(xs[rank].to(dev))
    ok = np.array_equal(col.cpu().numpy(), O.dist_collect([x.numpy() for x in xs]))
    res["dist_collect"] = bool(ok)
    res["ok"] = res["ok"] and bool(ok)

    # ---- sharded queue (BASELINE configs[3]): loss / prob / dq / ring contents vs the replicated oracle
    from moco_b200.NCE import ShardedMemoryMoCo
    Ns, C, K, T = 64, 128, 4096 * world, 0.07
    rng = np.random.default_rng(77)
    unit = lambda n: O.bf16_round(O.l2_normalize(rng.standard_normal((n, C)).astype(np.float32)))
    memory0 = unit(K)
    smod = ShardedMemoryMoCo(C, K, T)
    smod.memory.copy_(torch.from_numpy(memory0[smod.shard_row0:smod.shard_row0 + smod.shard_rows]))
    smod = smod.to(dev)
    orc = O.MemoryMoCoOracle(memory0, T)
    ok_s = True
    for step in range(3):
        q_all, k_all = unit(Ns * world), unit(Ns * world)          # every rank draws the same global batch
        own = slice(rank * Ns, (rank + 1) * Ns)
        pre = orc.memory.copy()
        out = orc.logits(q_all[own], k_all[own])
        ref_loss, ref_prob = O.nce_softmax_loss(out), O.prob_metric(out)
        ref_dq = O.nce_backward_dq(q_all[own], k_all[own], pre, T)
        orc.enqueue(k_all)
        qt = torch.from_numpy(q_all[own]).to(dev).requires_grad_(True)
        loss, prob = smod.forward_loss(qt, torch.from_numpy(k_all[own]).to(dev), torch.from_numpy(k_all).to(dev))
        loss.backward()
        e_l, e_p = abs(float(loss) - ref_loss), abs(float(prob) - ref_prob) / ref_prob
        e_dq = float(np.abs(qt.grad.cpu().numpy() - ref_dq).max() / np.abs(ref_dq).max())
        ok_s = ok_s and e_l < 2e-4 and e_p < 1e-3 and e_dq < 5e-3 and smod.index == orc.index
        res[f"sharded_step{step}"] = [e_l, e_p, e_dq]
    torch.cuda.synchronize()
    dist.barrier()                      # every rank's last enqueue has completed (a training loop has DDP's all-reduce here)
    ok_s = ok_s and np.array_equal(smod.full_memory().cpu().numpy(), orc.memory)
    if rank == 0:                       # the reference checkpoints on rank 0 only (train.py:226-228): no collective allowed
        sd = smod.state_dict()
        ok_s = ok_s and sorted(sd) == ["memory", "params"] and tuple(sd["memory"].shape) == (K, C) \
            and np.array_equal(sd["memory"].cpu().numpy(), orc.memory)
    dist.barrier()
    res["sharded_queue"] = bool(ok_s)
    res["ok"] = res["ok"] and bool(ok_s)
    res["max_err"] = {"loss_abs": max(res[f"sharded_step{i}"][0] for i in range(3)),
                      "prob_rel": max(res[f"sharded_step{i}"][1] for i in range(3)),
                      "dq_rel": max(res[f"sharded_step{i}"][2] for i in range(3)), "shufflebn": "bit-exact"}
    res["shufflebn"] = all(v for k, v in res.items() if k.startswith(("iter", "nhwc", "dist_collect")))
    flag = torch.tensor([1.0 if res["ok"] else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # parity must hold on EVERY rank
    res["ok_all_ranks"] = bool(flag.item() > 0.5)
    return res


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = correctness(rank, world, dev)
    res["ok"] = res["ok"] and res["ok_all_ranks"]

    # ---- bandwidth: BASELINE batch (256 x 3 x 224 x 224), fp32 and bf16, bulk-async vs LDG kernels
    ctx = ShuffleContext.get()
    n = 256
    for dt, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        x = torch.randn(n, 3, 224, 224, device=dev).to(dt)
        for flags, kname in ((_lib.GATHER_AUTO, "bulk"), (_lib.GATHER_LDG, "ldg")):
            ctx.gather_flags = flags
            for _ in range(3):
                DistributedShufle.forward_shuffle(x, 5)
            torch.cuda.synchronize()
            dist.barrier()
            iters = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                y, _ = DistributedShufle.forward_shuffle(x, 5)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            row_bytes = 3 * 224 * 224 * x.element_size()
            fwd, _ = O.get_shuffle_ids(n * world, 5)
            remote = int(((fwd[rank * n:(rank + 1) * n] // n) != rank).sum())
            res[f"fwd_shuffle_{tag}_{kname}_us"] = float(ms) * 1e3          # staging copy + barrier + gather
            res[f"fwd_shuffle_{tag}_{kname}_pull_GBps"] = n * row_bytes / (float(ms) * 1e-3) / 1e9
            res[f"remote_rows_{tag}"] = remote
            fwd_t = torch.from_numpy(fwd).to(dev)
            gathered = [torch.empty_like(x) for _ in range(world)]          # expected rows via NCCL, for the check
            dist.all_gather(gathered, x)
            exp = torch.cat(gathered)[fwd_t[rank * n:(rank + 1) * n]]
            okb = bool(torch.equal(y, exp))
            res[f"fwd_shuffle_{tag}_{kname}_ok"] = okb
            res["ok"] = res["ok"] and okb
        ctx.gather_flags = _lib.GATHER_AUTO
    # gather kernel alone (no staging copy / barrier): time the C call on pre-staged data
    x = torch.randn(n, 3, 224, 224, device=dev).bfloat16()
    buf = ctx._staging("fwd", x.numel() * 2)
    buf.tensor(x.shape, x.dtype).copy_(x)
    ctx.barrier()
    fwd, _ = O.get_shuffle_ids(n * world, 9)
    src = torch.from_numpy(fwd[rank * n:(rank + 1) * n].copy()).to(dev)
    out = torch.empty_like(x)
    lib = _lib.load()
    for flags, kname in ((0, "bulk"), (1, "ldg")):
        ctx.gather_flags = flags

        def call():          # the pull kernel exactly as forward_shuffle launches it (cross-GPU event folded in)
            return ctx._pull(buf.table, n, src, 3 * 224 * 224 * 2, out.data_ptr(), synced=True)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        remote = int(((fwd[rank * n:(rank + 1) * n] // n) != rank).sum())
        res[f"gather_only_bf16_{kname}_us"] = float(ms) * 1e3
        res[f"gather_only_bf16_{kname}_nvlink_GBps"] = remote * 3 * 224 * 224 * 2 / (float(ms) * 1e-3) / 1e9
    # ---- push vs pull over NVLink, 100 % remote: the same copy kernels with the roles of the pointers swapped
    #      (pull: source = the next rank's buffer, destination local; push: source local, destination = the next rank's)
    nxt = (rank + 1) % world
    big = ctx._staging("pushpull", x.numel() * 2)
    big2 = ctx._staging("pushpull", x.numel() * 2)
    ident = torch.arange(n, dtype=torch.long, device=dev)
    local_src = torch.randn(n, 3, 224, 224, device=dev).bfloat16()
    row_bytes = 3 * 224 * 224 * 2
    ctx.barrier()
    for flags, kname in ((0, "bulk"), (1, "ldg")):
        for mode in ("pull", "push"):
            if mode == "pull":
                table = (ctypes.c_void_p * 1)(big.ptrs[nxt])
                dst_ptr = out.data_ptr()
            else:
                table = (ctypes.c_void_p * 1)(local_src.data_ptr())
                dst_ptr = big2.ptrs[nxt]

            def call():
                return lib.moco_shuffle_gather(table, 1, n, ident.data_ptr(), n, row_bytes, dst_ptr, flags,
                                               torch.cuda.current_stream().cuda_stream)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[f"{mode}_{kname}_all_remote_GBps"] = n * row_bytes / (float(ms) * 1e-3) / 1e9
            dist.barrier()
    ctx.barrier()
    dist.barrier()
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
