"""the head alone (a7-a11: logits, loss, prob, dq; enqueue excluded) on HOST cores, two ways:
the numpy oracle (oracle/moco_oracle.py) and the reference's sequence of PyTorch CPU ops restated op for op.
Prints one JSON line per shape.  Runs anywhere (no GPU); states the core count with every number."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import moco_oracle as O  # noqa: E402

SHAPES = {"c1": (32, 128, 1024), "c2": (256, 128, 16384), "c3": (256, 128, 65536), "c5": (512, 256, 262144)}
T = 0.07


def torch_head(q, k, mem):
    q = q.clone().requires_grad_(True)
    l_pos = (q * k).sum(dim=-1, keepdim=True)
    l_neg = torch.mm(q, mem.clone().detach().t())
    out = (torch.cat((l_pos, l_neg), dim=1) / T).contiguous()
    loss = F.cross_entropy(out, torch.zeros(out.shape[0], dtype=torch.long))
    prob = F.softmax(out, dim=1)[:, 0].mean()
    loss.backward()
    return float(loss.detach()), float(prob.detach())


def main():
    names = sys.argv[1:] or list(SHAPES)
    rng = np.random.default_rng(0)
    for name in names:
        N, C, K = SHAPES[name]
        unit = lambda n: O.l2_normalize(rng.standard_normal((n, C)).astype(np.float32))
        q, k, mem = unit(N), unit(N), unit(K)
        reps = 30 if K <= 1024 else (5 if K <= 65536 else 1)

        def oracle():
            out = O.MemoryMoCoOracle(mem, T).logits(q, k)
            return O.nce_softmax_loss(out), O.prob_metric(out), O.nce_backward_dq(q, k, mem, T)
        oracle()
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle()
        t_or = (time.perf_counter() - t0) / reps
        tq, tk, tm = torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(mem)
        torch_head(tq, tk, tm)
        t0 = time.perf_counter()
        for _ in range(reps):
            torch_head(tq, tk, tm)
        t_th = (time.perf_counter() - t0) / reps
        print(json.dumps({"case": f"cpu_head_{name}", "N": N, "C": C, "K": K, "cores": os.cpu_count(),
                          "torch_threads": torch.get_num_threads(), "numpy_oracle_ms": round(t_or * 1e3, 2),
                          "torch_cpu_reference_ops_ms": round(t_th * 1e3, 2)}), flush=True)


if __name__ == "__main__":
    main()
