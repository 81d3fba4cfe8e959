"""Where do the stats kernel's roles wait?  (MOCO_DEBUG_MODE=8: per-CTA clock64 accumulators.)"""
import ctypes, json, os, sys
os.environ["MOCO_DEBUG_MODE"] = os.environ.get("MOCO_DEBUG_MODE", "8")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
from moco_b200 import _lib
lib = _lib.load()
N, C, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 256, 262144)
T = 0.07
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 4
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
n = 148 * 8
buf = (ctypes.c_ulonglong * n)()
assert lib.moco_debug_read_prof(wp, N, C, buf, n) == 0
a = np.array(buf[:], dtype=np.float64).reshape(148, 8)
a = a[a[:, 4] > 0]
names = ["mma_wait_tempty", "mma_wait_full", "producer_wait_empty", "epi_wait_tfull", "mma_thread_total", "epi_total"]
print(json.dumps({"ctas": int(a.shape[0]), "mode": os.environ["MOCO_DEBUG_MODE"], "flags": flags,
                  **{nm: round(float(a[:, i].mean())) for i, nm in enumerate(names)}}))
