"""ShardedMemoryMoCo -- MemoryMoCo with the queue partitioned across ranks (BASELINE configs[3]).

Same math as ``moco/NCE/Contrast.py`` (bl0/moco) for a queue of K rows, but rank r stores only ring
slots [r*K/W, (r+1)*K/W): 1/W of the memory and of the HBM bytes per step.  Every rank scores ALL W*N
queries of the step against its shard on the tcgen05 kernels; two small NCCL collectives stitch the
softmax: an all_gather of one (max, sum) pair per query and a reduce_scatter of the [W*N, C] partial
gradients.  The only key exchange is the all_gather of this step's keys (``k_all``, which ShuffleBN's
un-shuffle already produced).  See ``include/moco_b200.h`` (moco_nce_shard_*).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch import nn

from .. import _lib


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _ShardedNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, k_all, mod):
        lib = _lib.load()
        rank, world = _world()
        _lib.require_cuda(q, k, k_all, mod.memory)
        q_d = q.detach().contiguous()
        k_d = k.detach().to(q_d.dtype).contiguous()
        k_all = k_all.detach().to(q_d.dtype).contiguous()
        N, C = q_d.shape
        Nq = N * world
        if k_all.shape[0] != Nq:
            raise ValueError(f"ShardedMemoryMoCo: k_all has {k_all.shape[0]} rows, expected world*N = {Nq}")
        dev = q_d.device
        if world > 1:
            q_all = torch.empty(Nq, C, dtype=q_d.dtype, device=dev)
            dist.all_gather_into_tensor(q_all, q_d)                     # rank-major, like k_all
        else:
            q_all = q_d
        shard = mod._queue_bf16()
        Ks = shard.shape[0]
        ws, ws_ptr, ws_bytes = mod._workspace(Nq, C, dev)
        f32 = dict(dtype=torch.float32, device=dev)
        dt = _lib.dtype_code(q_all)
        stream = _lib.cur_stream()
        inv_T = 1.0 / mod.temperature
        ms = torch.empty(Nq, 2, **f32)
        # loss statistics AND the unnormalised gradient partials from one sweep over the shard when a gradient is
        # wanted and the temperature allows it (same policy as moco_nce_fwd, include/moco_b200.h)
        flags = mod.kernel_flags
        if (q.requires_grad and not (flags & (_lib.NCE_TWO_PASS | _lib.NCE_DQ_V1 | _lib.NCE_ONE_PASS))
                and inv_T <= _lib.ONE_PASS_MAX_INV_T):
            flags |= _lib.NCE_ONE_PASS
        if not q.requires_grad:
            flags &= ~_lib.NCE_ONE_PASS
        _lib.check(lib.moco_nce_shard_stats(q_all.data_ptr(), k_all.data_ptr(), dt, shard.data_ptr(), Nq, C, Ks, inv_T,
                                            ms.data_ptr(), ws_ptr, ws_bytes, flags, stream),
                   "moco_nce_shard_stats")
        if world > 1:
            ms_all = torch.empty(world, Nq, 2, **f32)
            dist.all_gather_into_tensor(ms_all, ms)
        else:
            ms_all = ms
        lse, loss_rows, prob_rows = (torch.empty(Nq, **f32) for _ in range(3))
        loss_prob_all = torch.empty(2, **f32)
        _lib.check(lib.moco_nce_shard_merge(ms_all.data_ptr(), world, Nq, C, inv_T, lse.data_ptr(), loss_rows.data_ptr(),
                                            prob_rows.data_ptr(), loss_prob_all.data_ptr(), ws_ptr, ws_bytes, stream),
                   "moco_nce_shard_merge")
        own = slice(rank * N, (rank + 1) * N)
        loss = loss_rows[own].mean()          # this rank's loss is the mean over ITS rows (train.py:263)
        prob = prob_rows[own].mean()
        ctx.dq = None
        if q.requires_grad:
            o_part = torch.empty(Nq, C, **f32)
            _lib.check(lib.moco_nce_shard_dq(q_all.data_ptr(), dt, shard.data_ptr(), lse.data_ptr(), Nq, C, Ks, inv_T,
                                             o_part.data_ptr(), ws_ptr, ws_bytes, flags, stream),
                       "moco_nce_shard_dq")
            if world > 1:
                o_own = torch.empty(N, C, **f32)
                dist.reduce_scatter_tensor(o_own, o_part)
            else:
                o_own = o_part
            dq = torch.empty(N, C, **f32)
            prob_own = prob_rows[own].contiguous()
            _lib.check(lib.moco_nce_shard_dq_finish(o_own.data_ptr(), k_d.data_ptr(), _lib.dtype_code(k_d),
                                                    prob_own.data_ptr(), N, C, inv_T, dq.data_ptr(), stream),
                       "moco_nce_shard_dq_finish")
            ctx.dq = dq
        ctx.q_dtype = q.dtype
        ctx.mark_non_differentiable(prob)
        return loss, prob

    @staticmethod
    def backward(ctx, g_loss, g_prob):
        if ctx.dq is None:
            return None, None, None, None
        return (ctx.dq * g_loss).to(ctx.q_dtype), None, None, None


class ShardedMemoryMoCo(nn.Module):
    """Queue of `queue_size` rows split evenly over the default process group ("block" layout: ring slot g
    lives on rank g // (K/W) at local row g % (K/W)).  `forward_loss(q, k, k_all) -> (loss, prob)`."""

    def __init__(self, feature_dim, queue_size, temperature=0.07):
        super().__init__()
        rank, world = _world()
        if queue_size % world != 0:
            raise ValueError(f"queue_size {queue_size} is not divisible by the world size {world}")
        self.queue_size = queue_size
        self.temperature = temperature
        self.index = 0
        self.kernel_flags = _lib.NCE_AUTO
        self.shard_rows = queue_size // world
        self.shard_row0 = rank * self.shard_rows
        self.register_buffer('params', torch.tensor([-1]))
        # identical initial queue on every rank (same global RNG stream as the reference, Contrast.py:16-17),
        # of which this rank keeps its block
        stdv = 1. / math.sqrt(feature_dim / 3)
        full = torch.rand(queue_size, feature_dim).mul_(2 * stdv).add_(-stdv)
        self.register_buffer('memory', full[self.shard_row0:self.shard_row0 + self.shard_rows].clone())
        self.register_buffer('memory_bf16', torch.empty(0, dtype=torch.bfloat16), persistent=False)
        self._bf16_src = None
        self._ws = {}

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._bf16_src = None
        self._ws = {}
        return out

    def _queue_bf16(self):
        mem = self.memory
        tag = (mem.data_ptr(), mem._version)
        if self._bf16_src != tag or self.memory_bf16.shape != mem.shape or self.memory_bf16.device != mem.device:
            self.memory_bf16 = torch.empty_like(mem, dtype=torch.bfloat16)
            lib = _lib.load()
            _lib.check(lib.moco_f32_to_bf16(mem.data_ptr(), self.memory_bf16.data_ptr(), mem.numel(), _lib.cur_stream()),
                       "moco_f32_to_bf16")
            self._bf16_src = tag
        return self.memory_bf16

    def _workspace(self, Nq, C, dev):
        key = (Nq, C, dev)
        hit = self._ws.get(key)
        if hit is None:
            lib = _lib.load()
            nbytes = int(lib.moco_nce_workspace_bytes(Nq, C, self.shard_rows))
            t = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            hit = self._ws[key] = (t, t.data_ptr() + (-t.data_ptr()) % 256, nbytes)
        return hit

    @torch.no_grad()
    def enqueue(self, k_all):
        lib = _lib.load()
        k_all = k_all.detach().contiguous()
        n_all, C = k_all.shape
        shard = self._queue_bf16()
        _lib.check(lib.moco_queue_enqueue_shard(shard.data_ptr(), self.memory.data_ptr(), k_all.data_ptr(),
                                                _lib.dtype_code(k_all), n_all, C, self.queue_size, self.index,
                                                self.shard_row0, self.shard_rows, _lib.cur_stream()),
                   "moco_queue_enqueue_shard")
        self._bf16_src = (self.memory.data_ptr(), self.memory._version)
        self.index = (self.index + n_all) % self.queue_size

    def forward_loss(self, q, k, k_all):
        loss, prob = _ShardedNCE.apply(q, k.detach(), k_all, self)
        self.enqueue(k_all)
        return loss, prob

    @torch.no_grad()
    def full_memory(self):
        """[K, C] fp32 queue gathered from every rank (checkpoint-compatible with MemoryMoCo's `memory`)."""
        rank, world = _world()
        if world == 1:
            return self.memory.clone()
        out = torch.empty(self.queue_size, self.memory.shape[1], dtype=self.memory.dtype, device=self.memory.device)
        dist.all_gather_into_tensor(out, self.memory.contiguous())
        return out
