"""Timeline of CTA 0 of the v1 stats kernel (MOCO_DEBUG_MODE=24): per-stage MMA-thread timestamps, producer
timestamps, epilogue timestamps.  Prints deltas in cycles."""
import ctypes, json, os, sys
os.environ["MOCO_DEBUG_MODE"] = "24"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
from moco_b200 import _lib
lib = _lib.load()
N, C, K, T = 512, 256, 262144, 0.07
flags = int(sys.argv[1]) if len(sys.argv) > 1 else (64 | 4)
dev = torch.device("cuda:0")
q = F.normalize(torch.randn(N, C, device=dev), dim=1).bfloat16()
k = F.normalize(torch.randn(N, C, device=dev), dim=1).bfloat16()
queue = F.normalize(torch.randn(K, C, device=dev), dim=1).bfloat16()
f32 = dict(dtype=torch.float32, device=dev)
lse, lr, pr, lp = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(2, **f32)
wsb = lib.moco_nce_workspace_bytes(N, C, K)
ws = torch.zeros(wsb + 256, dtype=torch.uint8, device=dev)
wp = ws.data_ptr() + (-ws.data_ptr()) % 256
for _ in range(3):
    rc = lib.moco_nce_fwd(q.data_ptr(), k.data_ptr(), 1, queue.data_ptr(), N, C, K, 1 / T, None, lse.data_ptr(), lr.data_ptr(),
                          pr.data_ptr(), lp.data_ptr(), None, wp, wsb, flags, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.moco_last_error()
torch.cuda.synchronize()
n = 8192
buf = (ctypes.c_ulonglong * n)()
assert lib.moco_debug_read_prof(wp, N, C, buf, n) == 0
a = np.array(buf[:], dtype=np.int64)
mma = a[4096:4096 + 220].reshape(-1, 2)[:108]
prod = a[5120:5120 + 108]
epi = a[6144:6144 + 81].reshape(-1, 3)[:27]
t0 = mma[0, 0]
print("MMA thread per stage: [after wait(full)] [after 4 MMAs + commit]  (cycles since first stage)")
for i in range(0, 24):
    print(i, mma[i, 0] - t0, mma[i, 1] - t0, " wait+gap:", mma[i, 0] - (mma[i - 1, 1] if i else t0), " issue:", mma[i, 1] - mma[i, 0],
          " producer got empty at", prod[i] - t0)
print("epilogue warp4 per tile: [tfull seen] [ld done + tempty arrive] [fold done]")
for i in range(0, 10):
    print(i, epi[i] - t0, " drain:", epi[i, 1] - epi[i, 0], " math:", epi[i, 2] - epi[i, 1])
print("steady-state stage period:", (mma[100, 0] - mma[20, 0]) / 80.0, "cycles (ideal 512)")
