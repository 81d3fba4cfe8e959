"""CPU restatement of one reference MoCo iteration (train.py:244-283 of bl0/moco) -- TEST/BASELINE
INFRASTRUCTURE ONLY (see oracle/moco_oracle.py header).  Used by bench.py's `cpu_baseline` leg and
`--impl reference` arm; the reference itself is Python and cannot travel to the GPU box, so this is the
"port" the bench times on the host cores.

Hot path = the numpy oracle (ShuffleBN permute, logits, InfoNCE, dq, enqueue); encoders / SGD / EMA are
host PyTorch on CPU in fp32, as in the reference (the encoder class is the same one the GPU arm uses).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import moco_oracle as O


class CpuMoCoStep:
    def __init__(self, arch: str, feat_dim: int, K: int, T: float, batch: int, lr: float = 0.03, seed: int = 0):
        from moco_b200 import encoders
        torch.manual_seed(seed)
        ctor = getattr(encoders, arch)
        self.model = ctor(low_dim=feat_dim)
        self.model_ema = ctor(low_dim=feat_dim)
        self.model_ema.load_state_dict(self.model.state_dict())          # moment_update(model, model_ema, 0), train.py:133
        self.model.train()
        self.model_ema.eval()
        for m in self.model_ema.modules():                               # set_bn_train, util.py:114-121
            if m.__class__.__name__.find("BatchNorm") != -1:
                m.train()
        stdv = O.queue_init_bound(feat_dim)
        memory = (torch.rand(K, feat_dim).mul_(2 * stdv).add_(-stdv)).numpy()   # Contrast.py:16-17
        self.contrast = O.MemoryMoCoOracle(memory, T)
        self.opt = torch.optim.SGD(self.model.parameters(), lr=lr * batch / 256, momentum=0.9, weight_decay=1e-4)
        self.alpha = 0.999
        self.batch = batch

    def step(self, inputs: torch.Tensor, epoch: int = 1):
        x1, x2 = torch.split(inputs, [3, 3], dim=1)                       # train.py:250
        feat_q = self.model(x1)                                           # :256
        with torch.no_grad():
            (x2s,), binds = O.forward_shuffle([x2.numpy()], epoch)        # :258 (world 1)
            feat_k = self.model_ema(torch.from_numpy(np.ascontiguousarray(x2s)))   # :259
            k_all, (k_loc,) = O.backward_shuffle([feat_k.numpy()], binds, True)    # :260
        q = feat_q.detach().numpy()
        pre = self.contrast.memory.copy()                                 # Contrast.py:25 clone
        out = self.contrast.forward(q, k_loc, k_all)                      # :262
        loss = O.nce_softmax_loss(out)                                    # :263
        prob = O.prob_metric(out)                                         # :264
        dq = O.nce_backward_dq(q, k_loc, pre, self.contrast.temperature)  # :273 (autograd restated)
        self.opt.zero_grad()
        feat_q.backward(torch.from_numpy(dq.astype(np.float32)))
        self.opt.step()                                                   # :274
        with torch.no_grad():                                             # :277
            for p, pe in zip(self.model.parameters(), self.model_ema.parameters()):
                pe.mul_(self.alpha).add_(p.detach(), alpha=1 - self.alpha)
        return loss, prob


def time_cpu_arm(arch: str, feat_dim: int, K: int, T: float, batch: int, steps: int, warmup: int,
                 budget_s: float = 150.0):
    """images/sec of the CPU step on a bounded sample (`batch` images per step)."""
    torch.set_num_threads(torch.get_num_threads())
    st = CpuMoCoStep(arch, feat_dim, K, T, batch)
    g = torch.Generator().manual_seed(1234)
    inputs = torch.randn(batch, 6, 224, 224, generator=g)
    t_w = time.time()
    for _ in range(max(1, warmup)):
        st.step(inputs)
    per = (time.time() - t_w) / max(1, warmup)
    steps = max(1, min(steps, int(budget_s / max(per, 1e-3))))
    t0 = time.time()
    for _ in range(steps):
        loss, prob = st.step(inputs)
    dt = time.time() - t0
    return {"images_per_s": batch * steps / dt, "ms_per_step": dt / steps * 1e3, "steps": steps,
            "batch": batch, "loss": loss, "threads": torch.get_num_threads()}
