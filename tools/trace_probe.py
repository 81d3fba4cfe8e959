"""Lab tool: per-tile timeline of the one-pass / dq kernel (CTA 0) from clock64 stamps.

Builds moco_b200/libmoco_b200_trace.so (-DMOCO_TRACE) HERE (needs nvcc), runs one gpu_lab shape on it and prints,
per tile, the cycle deltas of the MMA-issuing thread and of one softmax thread of each tile group.

    python tools/trace_probe.py build            # in the build container
    python tools/trace_probe.py run N C K [flags] # on the GPU box
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from moco_b200 import build as B
    if sys.argv[1] == "build":
        print(B.build_trace_variant())
        return
    import torch
    import torch.nn.functional as F
    N, C, K = (int(v) for v in sys.argv[2:5])
    flags = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    B.LIB = os.path.join(os.path.dirname(B.LIB), "libmoco_b200_trace.so")
    from moco_b200 import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(B.LIB)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    q = F.normalize(torch.randn(N, C, device=dev, generator=g), dim=1).bfloat16()
    k = F.normalize(torch.randn(N, C, device=dev, generator=g), dim=1).bfloat16()
    queue = F.normalize(torch.randn(K, C, device=dev, generator=g), dim=1).bfloat16()
    f32 = dict(dtype=torch.float32, device=dev)
    lse, lr, pr = (torch.zeros(N, **f32) for _ in range(3))
    lp, dq = torch.zeros(2, **f32), torch.zeros(N, C, **f32)
    wsb = lib.moco_nce_workspace_bytes(N, C, K)
    ws = torch.zeros(wsb + 256, dtype=torch.uint8, device=dev)
    wp = ws.data_ptr() + (-ws.data_ptr()) % 256
    if C <= 128:
        raw.moco_debug_evt_reset()
        raw.moco_debug_tail_evt_reset()
    for _ in range(8 if C <= 128 else 3):
        rc = lib.moco_nce_fwd(q.data_ptr(), k.data_ptr(), 1, queue.data_ptr(), N, C, K, 1 / 0.07, None, lse.data_ptr(),
                              lr.data_ptr(), pr.data_ptr(), lp.data_ptr(), dq.data_ptr(), wp, wsb, flags,
                              torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.moco_last_error()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (4 * 64 * 8))()
    # C <= 128 runs on nce_head128_sm100.cu, C in {192, 256} on nce_dq2_sm100.cu: each has its own trace buffer
    assert (raw.moco_debug_h128_trace if C <= 128 else raw.moco_debug_dq2_trace)(buf) == 0
    t = [[[buf[(r * 64 + i) * 8 + s] for s in range(8)] for i in range(64)] for r in range(4)]
    base = min(v for r in t for row in r for v in row if v > 0)
    names = {0: "MMA : waitP  gotP  PVissued commitKV | waitKV gotKV Sissued commitS",
             1: "SM g0: waitS gotS ld0done st0 allst stwait arrived", 2: "SM g1: (same)",
             3: ("KERNEL (C<=128): entry pdl_wait_done setup_done q_data q_staged o_full O_written all_done | row 1: S-issuer got q | row 2: rows converted, smem stores issued, fence.proxy.async done"
                 if C <= 128 else "KERNEL: entry setup_done q_staged o_full O_written all_done")}
    if C <= 128:
        evt = (ctypes.c_ulonglong * (64 * 4))()
        assert raw.moco_debug_evt(evt) == 0
        tev = (ctypes.c_ulonglong * (64 * 4))()
        assert raw.moco_debug_tail_evt(tev) == 0
        rows = [[evt[i * 4], evt[i * 4 + 1], tev[i * 4 + 2], tev[i * 4 + 3]] for i in range(8)]
        print("per launch (ns): sweep window | sweep last exit -> tail go | tail window | tail last exit -> next sweep first entry")
        for i in range(1, 8):
            r, nx = rows[i], rows[i + 1] if i + 1 < 8 else None
            if r[1] == 0:
                continue
            print(f"  launch {i}: {r[1] - r[0]:6d} | {r[2] - r[1]:6d} | {r[3] - r[2]:6d} | " + (f"{nx[0] - r[3]:6d}" if nx and nx[1] else "     -"))
        cta = (ctypes.c_ulonglong * (160 * 4))()
        assert raw.moco_debug_h128_cta(cta) == 0
        rows = [(cta[i * 4], cta[i * 4 + 1], cta[i * 4 + 2], cta[i * 4 + 3]) for i in range(160) if cta[i * 4]]
        t0 = min(r[0] for r in rows)
        starts = sorted(r[0] - t0 for r in rows)
        ends = sorted(r[1] - t0 for r in rows)
        durs = sorted(r[1] - r[0] for r in rows)
        print(f"CTAs {len(rows)}: entry skew (ns) min/median/max = {starts[0]}/{starts[len(starts) // 2]}/{starts[-1]}; "
              f"exit (ns after first entry) min/median/max = {ends[0]}/{ends[len(ends) // 2]}/{ends[-1]}; "
              f"per-CTA duration (ns) min/median/max = {durs[0]}/{durs[len(durs) // 2]}/{durs[-1]}")
        by_tiles = {}
        for r in rows:
            by_tiles.setdefault(int(r[3]), []).append(r[1] - r[0])
        print("  duration by tiles per CTA:", {k: (min(v), sorted(v)[len(v) // 2], max(v)) for k, v in sorted(by_tiles.items())})
    for r in range(4):
        print(names[r])
        for i in range(min(64, 24)):
            row = t[r][i]
            if not any(row):
                continue
            print(f"  tile {i:2d}: " + " ".join(f"{(v - base) if v else -1:7d}" for v in row))


if __name__ == "__main__":
    main()
