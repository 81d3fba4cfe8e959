"""BatchNormAct2d (csrc/bn_nhwc.cu) against torch's own ops on the ResNet-50 layer shapes at batch 256: accuracy vs an
fp32 evaluation of the same bf16 inputs, and CUDA-event times (L2 flushed between iterations) of forward and backward
next to ATen's batch_norm (+ add + relu) on the same tensors.

    python tools/bn_lab.py [iters] [batch]       # one JSON line per shape + a total line
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (H = W, C, relu, residual, how many such layers in ResNet-50)
R50 = [(112, 64, 1, 0, 1),
       (56, 64, 1, 0, 6), (56, 256, 1, 1, 3), (56, 256, 0, 0, 1),
       (56, 128, 1, 0, 1), (28, 128, 1, 0, 7), (28, 512, 1, 1, 4), (28, 512, 0, 0, 1),
       (28, 256, 1, 0, 1), (14, 256, 1, 0, 11), (14, 1024, 1, 1, 6), (14, 1024, 0, 0, 1),
       (14, 512, 1, 0, 1), (7, 512, 1, 0, 5), (7, 2048, 1, 1, 3), (7, 2048, 0, 0, 1)]


def timed(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot * 1e3 / iters


def frac_bad(a, b, rtol, atol):
    return float(((a - b).abs() > atol + rtol * b.abs()).float().mean())


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    from moco_b200.bn import BatchNormAct2d
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = {"ours_fwd_us": 0.0, "torch_fwd_us": 0.0, "ours_bwd_us": 0.0, "torch_bwd_us": 0.0}
    ok_all = True
    for hw, C, relu, has_res, count in R50:
        shape = (batch, C, hw, hw)
        x = (torch.randn(shape, device=dev, generator=g) * 1.7 + 0.3).bfloat16().contiguous(memory_format=torch.channels_last)
        res = torch.randn(shape, device=dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last) if has_res else None
        dy = torch.randn(shape, device=dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
        mod = BatchNormAct2d(C, relu=bool(relu)).to(dev)
        with torch.no_grad():
            mod.weight.copy_(torch.rand(C, device=dev, generator=g) + 0.5)
            mod.bias.copy_(torch.randn(C, device=dev, generator=g) * 0.2)
        ref = torch.nn.BatchNorm2d(C).to(dev)
        ref.load_state_dict(mod.state_dict())

        # ---- accuracy: fp32 evaluation of the same bf16 inputs
        xq = x.clone().requires_grad_(True)
        rq = res.clone().requires_grad_(True) if has_res else None
        y = mod(xq, rq)
        y.backward(dy)
        x32 = x.float().requires_grad_(True)
        r32 = res.float().requires_grad_(True) if has_res else None
        z = ref(x32)
        if has_res:
            z = z + r32
        if relu:
            # one ReLU mask (the kernels') on both sides: the sign of a pre-activation that is zero to rounding is
            # decided by the order of the fp32 operations, for whole groups of equal bf16-quantised inputs at once
            mask_diff = float(((y > 0) != (z.detach() > 0)).float().mean())
            z = z * (y > 0).float()
        else:
            mask_diff = 0.0
        z.backward(dy.float())
        M = batch * hw * hw
        sc = float(x32.grad.abs().max())
        errs = {
            "y_bad_frac": frac_bad(y.float(), z.detach(), 1 / 128, 2e-3),
            "dx_bad_frac": frac_bad(xq.grad.float(), x32.grad, 1 / 64, 4e-3 * sc),
            "dres_bad_frac": frac_bad(rq.grad.float(), r32.grad, 1 / 128, 1e-6) if has_res else 0.0,
            "dgamma_rel": float((mod.weight.grad - ref.weight.grad).abs().max() / ref.weight.grad.abs().max()),
            "dbeta_rel": float((mod.bias.grad - ref.bias.grad).abs().max() / ref.bias.grad.abs().max()),
            "running_mean_abs": float((mod.running_mean - ref.running_mean).abs().max()),
            "running_var_rel": float(((mod.running_var - ref.running_var).abs() / ref.running_var).max()),
            "nbt": int(mod.num_batches_tracked), "relu_mask_diff_frac": mask_diff,
        }
        ok = (errs["y_bad_frac"] < 1e-5 and errs["dx_bad_frac"] < 1e-4 and errs["dres_bad_frac"] < 1e-5
              and errs["dgamma_rel"] < 2e-3 and errs["dbeta_rel"] < 2e-3 and errs["running_mean_abs"] < 1e-5
              and errs["running_var_rel"] < 1e-4 and errs["nbt"] == 1 and mask_diff < 1e-4)
        ok_all = ok_all and ok

        # ---- timing: same tensors, ATen's bf16 channels_last path as the encoders used it before
        def ours_fwd():
            with torch.no_grad():
                mod(x, res)

        def torch_fwd():
            with torch.no_grad():
                t = ref(x)
                if has_res:
                    t = t + res
                if relu:
                    F.relu(t, inplace=True)

        xo = x.clone().requires_grad_(True)
        ro = res.clone().requires_grad_(True) if has_res else None
        yo = mod(xo, ro)
        xt = x.clone().requires_grad_(True)
        rt = res.clone().requires_grad_(True) if has_res else None
        t = ref(xt)
        if has_res:
            t = t + rt
        if relu:
            t = F.relu(t)

        def ours_bwd():
            torch.autograd.grad(yo, [xo, mod.weight, mod.bias] + ([ro] if has_res else []), dy, retain_graph=True)

        def torch_bwd():
            torch.autograd.grad(t, [xt, ref.weight, ref.bias] + ([rt] if has_res else []), dy, retain_graph=True)

        bytes_el = M * C * 2
        r = {"shape": f"{batch}x{C}x{hw}x{hw}", "relu": relu, "residual": has_res, "layers": count, "ok": ok, **errs,
             "ours_fwd_us": timed(ours_fwd, iters, flush), "torch_fwd_us": timed(torch_fwd, iters, flush),
             "ours_bwd_us": timed(ours_bwd, iters, flush), "torch_bwd_us": timed(torch_bwd, iters, flush)}
        r["fwd_GBps"] = bytes_el * (3 + has_res) / r["ours_fwd_us"] / 1e3
        r["bwd_GBps"] = bytes_el * (5 + (3 if has_res and relu else 0)) / r["ours_bwd_us"] / 1e3
        for k in tot:
            tot[k] += r[k] * count
        print(json.dumps(r))
        sys.stdout.flush()
        del x, res, dy, xq, rq, y, x32, r32, z, xo, ro, yo, xt, rt, t
        torch.cuda.empty_cache()
    print(json.dumps({"total_per_encoder_pass_us": tot, "ok": ok_all,
                      "note": "sum over the 53 BN layers of ResNet-50 at this batch; a MoCo step runs the forward twice "
                              "(query + key encoder) and the backward once"}))


if __name__ == "__main__":
    main()
