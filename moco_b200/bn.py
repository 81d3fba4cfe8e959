"""BatchNorm2d with the block's ReLU / residual add folded in, on this library's channels_last bf16 kernels.

The reference's encoders apply ``nn.BatchNorm2d`` -> [``out += residual``] -> [``nn.ReLU``] at
``moco/models/resnet.py:42-63,74-102,114,139-143,156-157``; ShuffleBN (``moco/util.py:69-93``) exists for exactly these
batch statistics.  Left to ATen's channels_last kernels they are 64 % of the GPU time of a step
(``profiles/r2_bench_launches_by_kernel.csv``).  :class:`BatchNormAct2d` is an ``nn.BatchNorm2d`` (same parameters,
buffers and ``state_dict`` keys, same running-statistics updates) whose training-mode forward / backward on CUDA
bf16 channels_last activations are two launches each of ``csrc/bn_nhwc.cu`` (``moco_bn_fwd_train`` / ``moco_bn_bwd``).
Everything else -- CPU tensors, eval mode, fp32 or NCHW activations, channel counts the kernels do not take -- runs
``nn.BatchNorm2d``'s own forward followed by the add and the ReLU, i.e. exactly what the reference does.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F

from . import _lib

_workspaces = {}
_enabled = True
# bench.py's live roofline pass: a list makes every fused call append (kind, algorithmic bytes, start event, end event)
_prof = None


def _timed(kind, nbytes, call):
    if _prof is None:
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = call()
    e1.record()
    _prof.append((kind, nbytes, e0, e1))
    return rc


def set_fused(flag: bool) -> None:
    """Process-wide switch (A/B timing, debugging): False sends every BatchNormAct2d through the torch ops."""
    global _enabled
    _enabled = bool(flag)


def _workspace(device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(_lib.load().moco_bn_workspace_bytes(), dtype=torch.uint8, device=device)   # zeroed once
        _workspaces[key] = ws
    return ws


def _rows_ok(t, like=None):
    return (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4 and t.numel() > 0
            and t.is_contiguous(memory_format=torch.channels_last) and (like is None or t.shape == like.shape))


class _BatchNormActFn(torch.autograd.Function):
    """y = relu?(batch_norm_train(x) [+ residual]); x, residual, y bf16 channels_last; weight / bias fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, num_batches_tracked, momentum, eps, relu):
        lib = _lib.load()
        N, C, H, W = x.shape
        M = N * H * W
        y = torch.empty_like(x)                                   # keeps the channels_last strides
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty_like(mean)
        ws = _workspace(x.device)
        # algorithmic bytes: statistics read x; apply reads x (+ residual) and writes y
        code = _timed("bn_fwd", M * C * 2 * (3 + (residual is not None)), lambda: lib.moco_bn_fwd_train(
            x.data_ptr(), residual.data_ptr() if residual is not None else None, y.data_ptr(), M, C,
            weight.data_ptr(), bias.data_ptr(),
            running_mean.data_ptr() if running_mean is not None else None,
            running_var.data_ptr() if running_var is not None else None,
            num_batches_tracked.data_ptr() if num_batches_tracked is not None else None,
            float(momentum), float(eps), int(relu), mean.data_ptr(), invstd.data_ptr(), ws.data_ptr(), ws.numel(),
            _lib.cur_stream()))
        _lib.check(code, "moco_bn_fwd_train")
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        # the ReLU mask of the backward is recomputed from x unless a residual went into it
        ctx.save_for_backward(x, y if (relu and residual is not None) else None, weight, bias, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, bias, mean, invstd = ctx.saved_tensors
        lib = _lib.load()
        N, C, H, W = x.shape
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        want_res = ctx.has_res and ctx.needs_input_grad[3]
        if want_res and not ctx.relu:
            dres, dres_ptr = dy, None                             # without a ReLU the residual's gradient is dy itself
        elif want_res:
            dres = torch.empty_like(x)
            dres_ptr = dres.data_ptr()
        else:
            dres, dres_ptr = None, None
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty_like(dgamma)
        ws = _workspace(x.device)
        # algorithmic bytes: reduce reads dy, x (+ y for the mask); apply reads the same and writes dx (+ d residual)
        ops = 2 * (2 + (y is not None)) + 1 + (dres_ptr is not None)
        code = _timed("bn_bwd", N * H * W * C * 2 * ops, lambda: lib.moco_bn_bwd(
            dy.data_ptr(), x.data_ptr(), y.data_ptr() if y is not None else None, N * H * W, C,
            weight.data_ptr(), bias.data_ptr(), mean.data_ptr(), invstd.data_ptr(), int(ctx.relu),
            int(ctx.has_res), dx.data_ptr(), dres_ptr, dgamma.data_ptr(), dbeta.data_ptr(),
            ws.data_ptr(), ws.numel(), _lib.cur_stream()))
        _lib.check(code, "moco_bn_bwd")
        return dx, dgamma, dbeta, dres, None, None, None, None, None, None


class BatchNormAct2d(nn.BatchNorm2d):
    """``nn.BatchNorm2d`` + optional residual add + optional ReLU (``forward(x, residual=None)``)."""

    def __init__(self, num_features, relu=False, **kw):
        super().__init__(num_features, **kw)
        self.relu = bool(relu)

    def _fusable(self, x, residual):
        C = self.num_features
        return (_enabled and self.training and self.affine and self.momentum is not None
                and _rows_ok(x) and (residual is None or _rows_ok(residual, x))
                and 64 <= C <= 2048 and (C & (C - 1)) == 0 and x.shape[1] == C
                and x.numel() // C > 1                 # a single value per channel: nn.BatchNorm2d's own error
                and self.weight.dtype == torch.float32 and self.weight.is_cuda
                and (self.running_mean is None or self.running_mean.dtype == torch.float32))

    def forward(self, x, residual=None):
        if self._fusable(x, residual):
            return _BatchNormActFn.apply(x, self.weight, self.bias, residual, self.running_mean, self.running_var,
                                         self.num_batches_tracked if self.track_running_stats else None,
                                         self.momentum, self.eps, self.relu)
        y = super().forward(x)
        if residual is not None:
            y = y + residual
        return F.relu(y, inplace=True) if self.relu else y

    def extra_repr(self):
        return super().extra_repr() + f", relu={self.relu}"


class _MaxPool3x3s2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        N, C, H, W = x.shape
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, C, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        taps = torch.empty((N, OH, OW, C), dtype=torch.uint8, device=x.device)
        _lib.check(lib.moco_maxpool3x3s2_fwd(x.data_ptr(), y.data_ptr(), taps.data_ptr(), N, H, W, C, _lib.cur_stream()),
                   "moco_maxpool3x3s2_fwd")
        ctx.save_for_backward(taps)
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (taps,) = ctx.saved_tensors
        N, C, H, W = ctx.shape
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=dy.device, memory_format=torch.channels_last)
        _lib.check(_lib.load().moco_maxpool3x3s2_bwd(dy.data_ptr(), taps.data_ptr(), dx.data_ptr(), N, H, W, C,
                                                     _lib.cur_stream()), "moco_maxpool3x3s2_bwd")
        return dx


class MaxPool3x3s2(nn.MaxPool2d):
    """``nn.MaxPool2d(kernel_size=3, stride=2, padding=1)`` (moco/models/resnet.py:119): this library's kernels for CUDA
    bf16 channels_last activations, ``nn.MaxPool2d``'s own forward for anything else."""

    def __init__(self):
        super().__init__(kernel_size=3, stride=2, padding=1)

    def forward(self, x):
        if _enabled and _rows_ok(x) and x.shape[1] % 8 == 0:
            return _MaxPool3x3s2Fn.apply(x)
        return super().forward(x)
