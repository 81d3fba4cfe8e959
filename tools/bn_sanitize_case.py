"""Small BatchNormAct2d / MaxPool3x3s2 forward + backward cases for compute-sanitizer runs:
    compute-sanitizer --tool racecheck python tools/bn_sanitize_case.py
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from moco_b200.bn import BatchNormAct2d, MaxPool3x3s2
    dev = torch.device("cuda:0")
    cl = torch.channels_last
    out = {}
    for N, C, H, relu, has_res in [(8, 64, 17, True, False), (4, 256, 14, True, True), (2, 2048, 7, True, True),
                                   (32, 64, 28, False, False), (3, 128, 9, True, False)]:
        mod = BatchNormAct2d(C, relu=relu).to(dev)
        x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_(True)
        res = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_(True) if has_res else None
        for _ in range(2):
            y = mod(x, res)
            y.backward(torch.randn_like(y))
        z = F.batch_norm(x.detach().float(), None, None, mod.weight.detach(), mod.bias.detach(), True, 0.1, mod.eps)
        if has_res:
            z = z + res.detach().float()
        if relu:
            z = F.relu(z)
        out[f"bn_{N}x{C}x{H}"] = float((y.float() - z).abs().max())
    pool = MaxPool3x3s2()
    for N, C, H, W in [(3, 64, 7, 7), (2, 64, 30, 31)]:
        x = F.relu(torch.randn(N, C, H, W, device=dev)).bfloat16().contiguous(memory_format=cl).requires_grad_(True)
        y = pool(x)
        y.backward(torch.randn_like(y))
        out[f"pool_{N}x{C}x{H}x{W}"] = float((y.float() - F.max_pool2d(x.detach().float(), 3, 2, 1)).abs().max())
    torch.cuda.synchronize()
    out["ok"] = all(v < 0.05 for v in out.values())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
