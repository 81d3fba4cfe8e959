"""BatchNormAct2d (csrc/bn_nhwc.cu through moco_bn_fwd_train / moco_bn_bwd) against torch's own BatchNorm2d -> add ->
ReLU (the reference's sequence, moco/models/resnet.py:42-63,74-102,156-157) evaluated in fp32 on the same bf16 inputs.
Tolerances: outputs within one bf16 ulp (rtol 2^-7) of the fp32 result; per-channel sums within 2e-3 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.bfloat16().contiguous(memory_format=torch.channels_last)


def _bad(a, b, rtol, atol):
    return float(((a - b).abs() > atol + rtol * b.abs()).float().mean())


def _pair(C, relu, dev, seed):
    from moco_b200.bn import BatchNormAct2d
    g = torch.Generator(device=dev).manual_seed(seed)
    mod = BatchNormAct2d(C, relu=relu).to(dev)
    with torch.no_grad():
        mod.weight.copy_(torch.rand(C, device=dev, generator=g) + 0.5)
        mod.bias.copy_(torch.randn(C, device=dev, generator=g) * 0.3)
        mod.running_mean.copy_(torch.randn(C, device=dev, generator=g))
        mod.running_var.copy_(torch.rand(C, device=dev, generator=g) + 0.5)
    ref = torch.nn.BatchNorm2d(C).to(dev)
    ref.load_state_dict(mod.state_dict())
    return mod, ref, g


@pytest.mark.parametrize("N,C,H,relu,has_res", [
    (8, 64, 17, True, False),        # M = 2312: ragged last pass, one slab
    (3, 128, 9, True, False),        # M = 243: fewer rows than one unrolled trip
    (4, 256, 14, True, True),        # relu(bn + residual): mask from y
    (4, 256, 14, False, True),       # bn + residual, no relu
    (2, 512, 7, False, False),       # downsample branch
    (2, 2048, 7, True, True),        # widest layer: 32 slabs
    (2, 64, 1, True, False),         # two rows
    (32, 64, 56, True, False),       # many row chunks per slab
])
def test_fwd_bwd_match_fp32_reference(N, C, H, relu, has_res):
    dev = torch.device("cuda:0")
    mod, ref, g = _pair(C, relu, dev, 1000 + C + H)
    shape = (N, C, H, H)
    x = _cl(torch.randn(shape, device=dev, generator=g) * 1.5 + 0.4)
    res = _cl(torch.randn(shape, device=dev, generator=g)) if has_res else None
    dy = _cl(torch.randn(shape, device=dev, generator=g))
    xq = x.clone().requires_grad_(True)
    rq = res.clone().requires_grad_(True) if has_res else None
    import moco_b200._lib as L
    before = L.launches
    y = mod(xq, rq)
    assert L.launches == before + 2, "the fused kernels did not run"
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    y.backward(dy)
    assert L.launches == before + 4

    x32 = x.float().requires_grad_(True)
    r32 = res.float().requires_grad_(True) if has_res else None
    z = ref(x32)
    if has_res:
        z = z + r32
    if relu:
        # The sign of a pre-activation that is zero to rounding (|z| ~ 1e-7; whole groups of elements share one such
        # value because the inputs are bf16-quantised) is decided by the order of the fp32 operations.  The forward
        # comparison below does not see it; the backward one uses one mask -- the kernels' -- on both sides.
        near0 = (z.detach().abs() < 1e-4)
        assert bool(((y > 0) == (z.detach() > 0))[~near0].all())
        z = z * (y > 0).float()
    z.backward(dy.float())
    assert _bad(y.float(), z.detach(), 1 / 128, 2e-3) < 1e-5
    if N * H * H > 1:
        sc = float(x32.grad.abs().max())
        assert _bad(xq.grad.float(), x32.grad, 1 / 64, 4e-3 * sc) < 2e-4
        assert float((mod.weight.grad - ref.weight.grad).abs().max()) < 2e-3 * float(ref.weight.grad.abs().max()) + 1e-4
    assert float((mod.bias.grad - ref.bias.grad).abs().max()) < 2e-3 * float(ref.bias.grad.abs().max()) + 1e-4
    if has_res:
        assert _bad(rq.grad.float(), r32.grad, 1 / 128, 1e-6) < 1e-5
    # running statistics and the step counter move exactly as nn.BatchNorm2d moves them
    assert float((mod.running_mean - ref.running_mean).abs().max()) < 1e-5
    if N * H * H > 1:
        assert float(((mod.running_var - ref.running_var).abs() / ref.running_var.abs()).max()) < 1e-4
    assert int(mod.num_batches_tracked) == int(ref.num_batches_tracked) == 1


def test_matches_atens_bf16_path_and_is_deterministic():
    """Against what the encoders ran before: ATen's batch_norm on the bf16 tensor itself, then add, then ReLU."""
    dev = torch.device("cuda:0")
    mod, ref, g = _pair(256, True, dev, 7)
    x = _cl(torch.randn(16, 256, 28, 28, device=dev, generator=g))
    res = _cl(torch.randn(16, 256, 28, 28, device=dev, generator=g))
    y1 = mod(x, res)
    t = F.relu(ref(x) + res)
    # ATen rounds the normalised value to bf16 BEFORE the add: where bn(x) and the residual cancel, its result carries
    # the absolute rounding error of the larger operand (2^-9 * |bn(x)|, |bn(x)| up to ~8 here)
    assert _bad(y1.float(), t.float(), 1 / 64, 4e-3) < 1e-3
    assert float((y1.float() - t.float()).abs().max()) < 0.07
    mod2, _, _ = _pair(256, True, dev, 7)
    y2 = mod2(x, res)
    assert torch.equal(y1, y2)
    assert torch.equal(mod.running_var, mod2.running_var)


def test_unfused_cases_run_the_reference_sequence():
    from moco_b200.bn import BatchNormAct2d
    import moco_b200._lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    # C = 96 is not a shape the kernels take; fp32 NCHW is not their layout; eval mode uses the running statistics
    for C, x, train in [(96, _cl(torch.randn(4, 96, 8, 8, device=dev, generator=g)), True),
                        (64, torch.randn(4, 64, 8, 8, device=dev, generator=g), True),
                        (64, _cl(torch.randn(4, 64, 8, 8, device=dev, generator=g)), False)]:
        mod = BatchNormAct2d(C, relu=True).to(dev).train(train)
        ref = torch.nn.BatchNorm2d(C).to(dev).train(train)
        before = L.launches
        y = mod(x, x)
        assert L.launches == before
        assert torch.equal(y, F.relu(ref(x) + x))


def test_resnet50_step_fused_is_as_close_to_fp32_as_the_torch_bf16_ops():
    """Whole encoder, forward + backward: fp32 torch ops are the truth; the fused BN kernels under bf16 autocast must be
    no further from it than ATen's own bf16 BatchNorm -> add -> ReLU sequence under the same autocast."""
    from moco_b200 import encoders, bn
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    models = [encoders.resnet50(128).to(dev).to(memory_format=torch.channels_last) for _ in range(3)]
    for m in models[1:]:
        m.load_state_dict(models[0].state_dict())
    x = torch.randn(16, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.linspace(-1, 1, 128, device=dev)
    outs = []
    for mod, mode in zip(models, ("fused", "torch_bf16", "fp32")):
        bn.set_fused(mode == "fused")
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode != "fp32"):
                q = mod(x)
            (q * w).sum().backward()
        finally:
            bn.set_fused(True)
        outs.append((q.detach().float(), mod.fc.weight.grad.float().clone(), mod.stem[0].weight.grad.float().clone(),
                     mod.stem[1].running_var.clone(), mod.layers[-1].bn3.running_mean.clone()))
    fused, torch_bf16, fp32 = outs
    for i, name in enumerate(("q", "fc.weight.grad", "stem conv weight.grad")):
        e_f = float((fused[i] - fp32[i]).norm() / fp32[i].norm())
        e_t = float((torch_bf16[i] - fp32[i]).norm() / fp32[i].norm())
        assert e_f < max(2.0 * e_t, 0.05), (name, e_f, e_t)
    assert float(((fused[3] - fp32[3]).abs() / fp32[3]).max()) < 2e-2          # first layer: same input on both sides
    e_f = float((fused[4] - fp32[4]).norm() / fp32[4].norm())
    e_t = float((torch_bf16[4] - fp32[4]).norm() / fp32[4].norm())
    assert e_f < max(2.0 * e_t, 0.05), ("last running_mean", e_f, e_t)


def test_cuda_graph_replay():
    from moco_b200.bn import BatchNormAct2d
    dev = torch.device("cuda:0")
    mod = BatchNormAct2d(128, relu=True).to(dev)
    x = _cl(torch.randn(8, 128, 14, 14, device=dev))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        mod(x)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            y = mod(x)
    torch.cuda.current_stream().wait_stream(s)
    x.copy_(_cl(torch.randn(8, 128, 14, 14, device=dev)))
    nbt = int(mod.num_batches_tracked)
    gr.replay()
    torch.cuda.synchronize()
    ref = torch.nn.BatchNorm2d(128).to(dev)
    z = F.relu(ref(x.float()))
    assert _bad(y.float(), z, 1 / 128, 2e-3) < 1e-5
    assert int(mod.num_batches_tracked) == nbt + 1


@pytest.mark.parametrize("N,C,H,W", [(4, 64, 112, 112), (3, 64, 7, 7), (2, 72, 9, 12), (1, 8, 1, 1), (2, 64, 2, 5)])
def test_maxpool3x3s2_matches_torch_bit_exact(N, C, H, W):
    """Post-ReLU data (windows full of equal zeros): same maxima, and the gradient goes to the FIRST maximum of each
    window, as torch.nn.functional.max_pool2d routes it (moco/models/resnet.py:119,158)."""
    from moco_b200.bn import MaxPool3x3s2
    import moco_b200._lib as L
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(N * 1000 + H)
    x = _cl(F.relu(torch.randn(N, C, H, W, device=dev, generator=g)))
    pool = MaxPool3x3s2()
    xq = x.clone().requires_grad_(True)
    before = L.launches
    y = pool(xq)
    assert L.launches == before + 1
    x32 = x.float().requires_grad_(True)
    z = F.max_pool2d(x32, 3, 2, 1)
    assert y.shape == z.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y.float(), z)
    dy = _cl(torch.randn(z.shape, device=dev, generator=g))
    y.backward(dy)
    z.backward(dy.float())
    ref = x32.grad.bfloat16()
    # same winner everywhere; where two or more windows route into one pixel the fp32 sums may be added in another order
    assert float((xq.grad != ref).float().mean()) < 1e-4
    assert bool(((xq.grad.float() - ref.float()).abs() <= ref.float().abs() / 64 + 1e-6).all())
    assert torch.equal(xq.grad == 0, ref == 0)
    # fp32 / NCHW input: nn.MaxPool2d's own path
    before = L.launches
    assert torch.equal(pool(x.float()), z.detach())
    assert L.launches == before


def _s2d_reference(x):
    """moco_crop_s2d_bf16's layout from torch ops: [N, 3, H, W] -> bf16 [N, 16, H/2+3, W/2+3]."""
    N, C, H, W = x.shape
    xp = F.pad(x.to(torch.bfloat16), (4, 2, 4, 2))
    R, Q = H // 2 + 3, W // 2 + 3
    xs = xp.view(N, C, R, 2, Q, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 12, R, Q)
    return F.pad(xs, (0, 0, 0, 0, 0, 4))


@pytest.mark.parametrize("src_dtype", [torch.float32, torch.bfloat16])
def test_crop_to_s2d_is_the_layout_the_header_defines(src_dtype):
    from moco_b200.util import DistributedShufle, crop_to_s2d_bf16
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    six = torch.randn(6, 6, 224, 224, device=dev, generator=g).to(src_dtype)
    crop = six[:, 3:]                                             # a channel slice of the 6-channel batch, read in place
    out = crop_to_s2d_bf16(crop)
    assert out.shape == (6, 16, 115, 115) and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out, _s2d_reference(crop))
    small = torch.randn(3, 3, 10, 6, device=dev, generator=g).to(src_dtype)
    assert torch.equal(crop_to_s2d_bf16(small), _s2d_reference(small))
    # single-GPU ShuffleBN: the permutation rides in the same kernel
    shuf, binds = DistributedShufle.forward_shuffle(crop, 3, cast_dtype=torch.bfloat16, channels_last="s2d")
    finds, _ = DistributedShufle.get_shuffle_ids(6, 3, dev)
    assert torch.equal(shuf, _s2d_reference(crop)[finds])


def test_stem_conv_on_s2d_input_is_the_7x7_convolution():
    """moco/models/resnet.py:112 on the plain NHWC crop vs the 4x4 convolution on the space-to-depth crop: same
    function of the same 7x7 parameter, forward and weight gradient (bf16 autocast, cuDNN both times)."""
    from moco_b200 import encoders
    from moco_b200.util import crop_to_channels_last_bf16, crop_to_s2d_bf16
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    stem = encoders.StemConv().to(dev).to(memory_format=torch.channels_last)
    x = torch.randn(8, 3, 224, 224, device=dev)
    dy = torch.randn(8, 64, 112, 112, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a = stem(crop_to_channels_last_bf16(x))
        b = stem(crop_to_s2d_bf16(x))
    assert a.shape == b.shape == (8, 64, 112, 112)
    ref = F.conv2d(x.bfloat16().float(), stem.weight.detach().bfloat16().float(), None, 2, 3)
    ea = float((a.float() - ref).abs().max())
    eb = float((b.float() - ref).abs().max())
    assert eb < max(2 * ea, 0.05), (ea, eb)
    a.backward(dy)
    ga = stem.weight.grad.clone()
    stem.weight.grad = None
    b.backward(dy)
    gb = stem.weight.grad
    assert float((ga - gb).norm() / ga.norm()) < 2e-2
