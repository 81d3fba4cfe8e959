"""CPU oracle for the encoder-side ops this library also runs (TEST INFRASTRUCTURE ONLY -- see moco_oracle.py's
header for who may import ``oracle/``): the BatchNorm -> [+= residual] -> [ReLU] groups, the stem max pooling and
the space-to-depth form of the first convolution, restated in plain numpy.

The reference delegates all three to PyTorch modules (``nn.BatchNorm2d``, ``nn.MaxPool2d``, ``nn.Conv2d``;
third-party, ``pytorch>=1.3``, reference ``README.md:14``), at these call sites (paths relative to /root/reference):
    moco/models/resnet.py:42-63    BasicBlock: bn1 -> relu, bn2 -> += residual -> relu
    moco/models/resnet.py:74-102   Bottleneck: bn1 -> relu, bn2 -> relu, bn3 -> += residual -> relu
    moco/models/resnet.py:112-119  conv1 (7x7 / 2 / pad 3), bn1, relu, maxpool (3x3 / 2 / pad 1)
    moco/models/resnet.py:139-143  downsample: conv1x1 -> bn
    moco/models/resnet.py:155-158  forward: conv1 -> bn1 -> relu -> maxpool
Pinned by ``tests/golden/encoder_ops.npz`` -- tensors captured INSIDE the reference's own ``ResNet`` / ``Bottleneck``
modules (forward values and autograd gradients) by ``tests/golden/gen_golden.py`` -- in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def batchnorm_stats(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Per-channel mean and BIASED variance over (N, H, W) of x [N, C, H, W] (nn.BatchNorm2d, training mode)."""
    x64 = x.astype(np.float64)
    return x64.mean(axis=(0, 2, 3)), x64.var(axis=(0, 2, 3))


def running_stats_update(running_mean, running_var, mean, var_biased, count: int, momentum: float = 0.1):
    """nn.BatchNorm2d's buffers after one training step: the running variance takes the UNBIASED batch variance."""
    unbiased = var_biased * (count / (count - 1.0))
    return ((1 - momentum) * running_mean + momentum * mean).astype(np.float32), \
           ((1 - momentum) * running_var + momentum * unbiased).astype(np.float32)


def bn_act_forward(x, gamma, beta, residual: Optional[np.ndarray] = None, relu: bool = False, eps: float = 1e-5):
    """resnet.py:96-102 (and :88-89, :156-157 without the residual): y = relu?(bn(x) [+ residual]).
    Returns (y, mean, invstd)."""
    mean, var = batchnorm_stats(x)
    invstd = 1.0 / np.sqrt(var + eps)
    z = (x.astype(np.float64) - mean[None, :, None, None]) * invstd[None, :, None, None] * gamma[None, :, None, None] \
        + beta[None, :, None, None]
    if residual is not None:
        z = z + residual
    if relu:
        z = np.maximum(z, 0.0)
    return z.astype(np.float32), mean, invstd


def bn_act_backward(x, gamma, beta, dy, residual: Optional[np.ndarray] = None, relu: bool = False, eps: float = 1e-5):
    """Autograd of bn_act_forward: returns (dx, dgamma, dbeta, dresidual-or-None).
    g = dy where the ReLU passed (y > 0), dbeta = sum g, dgamma = sum g * xhat,
    dx = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat))."""
    y, mean, invstd = bn_act_forward(x, gamma, beta, residual, relu, eps)
    g = dy.astype(np.float64)
    if relu:
        g = g * (y > 0)
    xhat = (x.astype(np.float64) - mean[None, :, None, None]) * invstd[None, :, None, None]
    m = x.shape[0] * x.shape[2] * x.shape[3]
    dbeta = g.sum(axis=(0, 2, 3))
    dgamma = (g * xhat).sum(axis=(0, 2, 3))
    dx = (gamma * invstd)[None, :, None, None] * (g - dbeta[None, :, None, None] / m - xhat * dgamma[None, :, None, None] / m)
    return dx.astype(np.float32), dgamma.astype(np.float32), dbeta.astype(np.float32), \
        (g.astype(np.float32) if residual is not None else None)


def maxpool3x3s2_forward(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """resnet.py:119,158: nn.MaxPool2d(3, 2, 1).  Returns (y, taps): taps[n, c, oh, ow] = kh * 3 + kw of the FIRST
    maximum in kh-then-kw scan order among the in-image taps (torch: `val > max || isnan(val)` from max = -inf)."""
    N, C, H, W = x.shape
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = np.full((N, C, OH, OW), -np.inf, dtype=x.dtype)
    taps = np.zeros((N, C, OH, OW), dtype=np.uint8)
    first = np.ones((OH, OW), dtype=bool)
    for kh in range(3):
        for kw in range(3):
            for oh in range(OH):
                ih = 2 * oh - 1 + kh
                if ih < 0 or ih >= H:
                    continue
                for ow in range(OW):
                    iw = 2 * ow - 1 + kw
                    if iw < 0 or iw >= W:
                        continue
                    v = x[:, :, ih, iw]
                    take = (v > y[:, :, oh, ow]) | np.isnan(v) | first[oh, ow]
                    y[:, :, oh, ow] = np.where(take, v, y[:, :, oh, ow])
                    taps[:, :, oh, ow] = np.where(take, kh * 3 + kw, taps[:, :, oh, ow])
                    first[oh, ow] = False
    return y, taps


def maxpool3x3s2_backward(dy: np.ndarray, taps: np.ndarray, in_shape) -> np.ndarray:
    """Every output gradient goes to the one input element that won its window."""
    N, C, H, W = in_shape
    dx = np.zeros(in_shape, dtype=np.float64)
    OH, OW = dy.shape[2], dy.shape[3]
    for oh in range(OH):
        for ow in range(OW):
            t = taps[:, :, oh, ow].astype(np.int64)
            ih, iw = 2 * oh - 1 + t // 3, 2 * ow - 1 + t % 3
            n_idx, c_idx = np.meshgrid(np.arange(N), np.arange(C), indexing="ij")
            np.add.at(dx, (n_idx, c_idx, ih, iw), dy[:, :, oh, ow])
    return dx.astype(np.float32)


def s2d_layout(x: np.ndarray) -> np.ndarray:
    """[N, 3, H, W] -> [N, 16, H/2 + 3, W/2 + 3] (moco_crop_s2d_bf16, include/moco_b200.h):
    out[n, (b*2+d)*3 + c, R, Q] = x[n, c, 2(R-2)+b, 2(Q-2)+d], zero outside the image, channels 12..15 zero."""
    N, C, H, W = x.shape
    R, Q = H // 2 + 3, W // 2 + 3
    out = np.zeros((N, 16, R, Q), dtype=x.dtype)
    for b in range(2):
        for d in range(2):
            for c in range(C):
                out[:, (b * 2 + d) * 3 + c, 2:2 + H // 2, 2:2 + W // 2] = x[:, c, b::2, d::2]
    return out


def stem_weight_s2d(w: np.ndarray) -> np.ndarray:
    """[64, 3, 7, 7] (resnet.py:112) -> [64, 16, 4, 4]: w'[o, (b*2+d)*3+c, a, e] = w[o, c, 2a+b-1, 2e+d-1] (tap -1 = 0),
    the kernel with which a 4x4 / stride 1 / pad 0 convolution over s2d_layout(x) equals the 7x7 / 2 / pad 3 one over x."""
    O = w.shape[0]
    out = np.zeros((O, 16, 4, 4), dtype=w.dtype)
    for a in range(4):
        for b in range(2):
            kh = 2 * a + b - 1
            if kh < 0:
                continue
            for e in range(4):
                for d in range(2):
                    kw = 2 * e + d - 1
                    if kw < 0:
                        continue
                    out[:, (b * 2 + d) * 3:(b * 2 + d) * 3 + 3, a, e] = w[:, :, kh, kw]
    return out


def conv2d_valid(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Stride-1, no-padding cross-correlation (what the stem runs on the s2d input); float64 accumulation."""
    N, C, H, W = x.shape
    O, _, KH, KW = w.shape
    out = np.zeros((N, O, H - KH + 1, W - KW + 1), dtype=np.float64)
    for kh in range(KH):
        for kw in range(KW):
            out += np.einsum("nchw,oc->nohw", x[:, :, kh:kh + out.shape[2], kw:kw + out.shape[3]].astype(np.float64),
                             w[:, :, kh, kw].astype(np.float64))
    return out.astype(np.float32)
