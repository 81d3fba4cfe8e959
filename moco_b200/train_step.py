"""One MoCo pretraining iteration -- the body of ``train_moco`` (train.py:244-283 of bl0/moco)
with the contrastive hot path on the sm_100a kernels.

Differences from the reference loop body, none of which change the math:
* the ShuffleBN image permute runs on a side stream, overlapped with the query-encoder forward;
* ``contrast.forward_loss`` replaces train.py:262-264 (logits never materialised);
* no per-step ``.item()`` host syncs (train.py:280-281): loss / prob stay on the device;
* encoders run under bf16 autocast in channels_last (the reference used Apex AMP, train.py:189-196);
* with ``channels_last=True`` the two crops are taken straight from the 6-channel batch (train.py:250-254) as
  bf16 NHWC by one kernel each (x1: ``crop_to_channels_last_bf16``; x2: inside the ShuffleBN publish), which is
  what autocast + cuDNN would have produced with two more passes over the images.
"""
from __future__ import annotations

import torch

from .NCE import MemoryMoCo
from .util import DistributedShufle, crop_to_channels_last_bf16, moment_update, set_bn_train


class MoCoStep:
    def __init__(self, model, model_ema, contrast: MemoryMoCo, optimizer, alpha: float = 0.999,
                 amp_dtype=torch.bfloat16, overlap_shuffle: bool = True, channels_last: bool = False):
        self.model, self.model_ema, self.contrast, self.optimizer = model, model_ema, contrast, optimizer
        self.alpha = alpha
        self.amp_dtype = amp_dtype
        # fused input path only where it is value-preserving: bf16 autocast would round the images identically
        self.nhwc = bool(channels_last) and amp_dtype is torch.bfloat16
        self.side = torch.cuda.Stream() if overlap_shuffle else None
        self.model.train()
        set_bn_train(self.model_ema)                     # train.py:235-236

    def _ema_module(self):
        return self.model.module if hasattr(self.model, "module") else self.model

    def __call__(self, x1: torch.Tensor, x2: torch.Tensor, epoch: int):
        """x1, x2: [N, 3, 224, 224] CUDA tensors (the two crops, train.py:250-254).
        Returns (loss, prob) as 0-d CUDA tensors."""
        main = torch.cuda.current_stream()
        if self.side is not None:
            # ShuffleBN forward (train.py:258) on the side stream while the query encoder runs
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side), torch.no_grad():
                x2_shuffled, backward_inds = DistributedShufle.forward_shuffle(x2, epoch, cast_dtype=self.amp_dtype,
                                                                               channels_last=self.nhwc)
            x2.record_stream(self.side)
        if self.nhwc:
            x1 = crop_to_channels_last_bf16(x1)
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            feat_q = self.model(x1)                                                  # train.py:256
        with torch.no_grad():
            if self.side is not None:
                main.wait_stream(self.side)
                x2_shuffled.record_stream(main)
            else:
                x2_shuffled, backward_inds = DistributedShufle.forward_shuffle(x2, epoch, cast_dtype=self.amp_dtype,
                                                                               channels_last=self.nhwc)
            with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
                feat_k = self.model_ema(x2_shuffled)                                 # train.py:259
            feat_k_all, feat_k = DistributedShufle.backward_shuffle(feat_k, backward_inds, return_local=True)
        loss, prob = self.contrast.forward_loss(feat_q, feat_k, feat_k_all)          # train.py:262-264
        self.optimizer.zero_grad(set_to_none=True)                                   # train.py:267-268
        loss.backward()                                                              # train.py:273
        self.optimizer.step()                                                        # train.py:274
        moment_update(self._ema_module(), self.model_ema, self.alpha)                # train.py:277
        return loss, prob
