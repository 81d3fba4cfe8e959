import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from moco_b200.NCE import ShardedMemoryMoCo
from oracle import moco_oracle as O
rng=np.random.default_rng(1); C,K,T,N=128,4096,0.07,64
unit=lambda n: O.bf16_round(O.l2_normalize(rng.standard_normal((n,C)).astype(np.float32)))
mem=unit(K); m=ShardedMemoryMoCo(C,K,T); m.memory.copy_(torch.from_numpy(mem)); m=m.cuda(); orc=O.MemoryMoCoOracle(mem,T)
for s in range(2):
    q,k=unit(N),unit(N); pre=orc.memory.copy(); out=orc.logits(q,k); dq=O.nce_backward_dq(q,k,pre,T); orc.enqueue(k)
    qt=torch.from_numpy(q).cuda().requires_grad_(True); l,p=m.forward_loss(qt,torch.from_numpy(k).cuda(),torch.from_numpy(k).cuda()); l.backward()
    print(s, abs(float(l)-O.nce_softmax_loss(out)), abs(float(p)-O.prob_metric(out)), np.abs(qt.grad.cpu().numpy()-dq).max()/np.abs(dq).max())
print('mem ok', np.array_equal(m.full_memory().cpu().numpy(), orc.memory))
