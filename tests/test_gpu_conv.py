"""-m gpu: every conv / deconv variant of the tcgen05 engine through the C ABI against a plain PyTorch fp32
reference of the same op evaluated on the SAME bf16-rounded operands, so the only differences are fp32
accumulation order and the bf16 rounding of the stored result (and of v^2 inside the fused GDN).

Tolerance: |err| <= 1.2e-2 * max|ref| per tensor (bf16 has 8 mantissa bits: 2^-8 = 3.9e-3 relative per
rounding, two roundings on the GDN path), relative RMS <= 4e-3."""
import pytest
import torch
import torch.nn.functional as F

from licos_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.bfloat16().float()


def _ref(x, w, bias, kind):
    if kind == _lib.CONV_5X5_S2:
        return F.conv2d(x, w, bias, stride=2, padding=2)
    if kind == _lib.CONV_3X3_S1:
        return F.conv2d(x, w, bias, stride=1, padding=1)
    return F.conv_transpose2d(x, w, bias, stride=2, padding=2, output_padding=1)


def _gdn_ref(v, beta, gamma, inverse):
    # what the kernel computes: gamma (bf16) x bf16(v^2), fp32 accumulate
    C = v.size(1)
    norm = F.conv2d(_bf(v * v), _bf(gamma).reshape(C, C, 1, 1), beta)
    return v * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))


def _check(got, ref, what, out_bf16):
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    rms = (err.pow(2).mean().sqrt() / (ref.pow(2).mean().sqrt() + 1e-12)).item()
    worst = err.max().item() / scale
    idx = tuple(int(i) for i in torch.nonzero(err == err.max())[0])
    print(f"{what}: max|err|/max|ref| = {worst:.3e}  rel-rms = {rms:.3e}  at {idx} got {got[idx].item():.5f} "
          f"ref {ref[idx].item():.5f}")
    tol = 1.2e-2 if out_bf16 else 2e-3
    assert worst <= tol and rms <= (4e-3 if out_bf16 else 1e-3), what


def _run(cuda, kind, B, cin, cout, H, W, epi=_lib.EPI_NONE, out_nchw=False, first=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = _bf(torch.randn(B, cin, H, W, generator=g))
    kk = 3 if kind == _lib.CONV_3X3_S1 else 5
    wshape = (cin, cout, kk, kk) if kind == _lib.DECONV_5X5_S2 else (cout, cin, kk, kk)
    w = _bf(torch.randn(wshape, generator=g) / (cin * kk * kk) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    ref = _ref(x, w, bias, kind)
    beta = gamma = None
    if epi in (_lib.EPI_GDN, _lib.EPI_IGDN):
        beta = torch.rand(cout, generator=g) + 0.5
        gamma = torch.rand(cout, cout, generator=g) * 0.05 + 0.1 * torch.eye(cout)
        ref = _gdn_ref(ref, beta, gamma, epi == _lib.EPI_IGDN)
    elif epi == _lib.EPI_RELU:
        ref = torch.relu(ref)

    in_layout = _lib.LAYOUT_NCHW_F32 if first else _lib.LAYOUT_NHWC_BF16
    out_layout = _lib.LAYOUT_NCHW_F32 if out_nchw else _lib.LAYOUT_NHWC_BF16
    xd = x.to(cuda) if first else x.permute(0, 2, 3, 1).contiguous().bfloat16().to(cuda)
    packed = ops.pack_conv_weight(w.to(cuda), kind, cout, cin, in_layout)
    bh = gh = None
    if beta is not None:
        bh, gh = ops.gdn_pack(beta.sqrt().to(cuda), gamma.sqrt().to(cuda), 0.0, 0.0, 0.0)  # hat = (.)^2
    out = ops.conv_forward(xd, kind=kind, epilogue=epi, in_layout=in_layout, out_layout=out_layout, in_c=cin,
                           out_c=cout, weight=packed, bias=bias.to(cuda), beta=bh, gamma=gh)
    torch.cuda.synchronize()
    got = out.cpu() if out_nchw else out.float().cpu().permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    _check(got, ref, f"kind={kind} {cin}->{cout} {H}x{W} epi={epi} nchw_out={out_nchw} first={first}", not out_nchw)
    if beta is not None:
        # training's variant: same launch also writes v = conv + bias (bf16 NHWC).  The output must not change, every
        # element of v must be written, and v must be what the EPI_NONE launch of the same conv stores.
        Bo, Co, OH, OW = ref.shape
        v = torch.full((Bo, OH, OW, Co), float("nan"), dtype=torch.bfloat16, device=cuda)
        try:
            out2 = ops.conv_forward(xd, kind=kind, epilogue=epi, in_layout=in_layout, out_layout=out_layout, in_c=cin,
                                    out_c=cout, weight=packed, bias=bias.to(cuda), beta=bh, gamma=gh, pre_act=v)
        except NotImplementedError:
            # only the generic first-layer kernel (conv_edge.cuh) refuses; FusedSequential then runs conv and GDN apart
            assert first and (cin not in (1, 3) or (W * 4) % 16 != 0)
            return
        plain = ops.conv_forward(xd, kind=kind, epilogue=_lib.EPI_NONE, in_layout=in_layout,
                                 out_layout=_lib.LAYOUT_NHWC_BF16, in_c=cin, out_c=cout, weight=packed, bias=bias.to(cuda))
        torch.cuda.synchronize()
        assert torch.equal(out2, out), "pre_act changed the layer's output"
        assert torch.equal(v.view(torch.int16), plain.view(torch.int16)), "pre_act differs from the conv's own output"


def test_conv3x3_s1_nhwc(cuda):
    _run(cuda, _lib.CONV_3X3_S1, 2, 64, 64, 16, 16)


def test_conv3x3_s1_relu_to_nchw_320(cuda):
    _run(cuda, _lib.CONV_3X3_S1, 1, 128, 320, 8, 24, epi=_lib.EPI_RELU, out_nchw=True)


def test_conv5x5_s2_plain(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 2, 128, 128, 32, 32)


def test_conv5x5_s2_gdn(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 2, 128, 128, 32, 32, epi=_lib.EPI_GDN)


def test_conv5x5_s2_gdn_192(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 1, 192, 192, 16, 48, epi=_lib.EPI_GDN)


def test_conv5x5_s2_relu(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 1, 128, 128, 32, 32, epi=_lib.EPI_RELU)


def test_conv5x5_s2_to_nchw_192(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 2, 128, 192, 32, 32, out_nchw=True)


def test_conv5x5_s2_odd_sizes(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 1, 64, 64, 35, 21, epi=_lib.EPI_GDN)
    _run(cuda, _lib.CONV_5X5_S2, 1, 64, 128, 2, 2, out_nchw=True)


def test_first_layer_rgb_gdn(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 2, 3, 128, 64, 64, epi=_lib.EPI_GDN, first=True)


def test_first_layer_single_band_and_13_bands(cuda):
    _run(cuda, _lib.CONV_5X5_S2, 1, 1, 128, 48, 40, epi=_lib.EPI_GDN, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 1, 13, 128, 33, 31, epi=_lib.EPI_RELU, first=True)


def test_deconv_igdn(cuda):
    _run(cuda, _lib.DECONV_5X5_S2, 2, 192, 128, 16, 16, epi=_lib.EPI_IGDN)


def test_deconv_plain_and_relu(cuda):
    _run(cuda, _lib.DECONV_5X5_S2, 1, 128, 128, 20, 9)
    _run(cuda, _lib.DECONV_5X5_S2, 1, 64, 64, 1, 1, epi=_lib.EPI_RELU)


def test_deconv_last_layer_narrow(cuda):
    for cout in (3, 1, 13):
        _run(cuda, _lib.DECONV_5X5_S2, 2, 128, cout, 32, 32, out_nchw=True)


def test_edge_kernels_ragged_and_relu(cuda):
    # fused-im2col first layer and GEMM+gather last layer on sizes that are not tile multiples
    _run(cuda, _lib.CONV_5X5_S2, 2, 3, 128, 33, 31, epi=_lib.EPI_GDN, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 1, 1, 192, 70, 18, epi=_lib.EPI_GDN, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 1, 3, 128, 256, 256, epi=_lib.EPI_NONE, first=True)
    _run(cuda, _lib.DECONV_5X5_S2, 2, 128, 3, 20, 9, out_nchw=True)
    _run(cuda, _lib.DECONV_5X5_S2, 1, 192, 1, 7, 45, epi=_lib.EPI_RELU, out_nchw=True)
    _run(cuda, _lib.DECONV_5X5_S2, 3, 128, 3, 128, 128, out_nchw=True)


def test_deconv_to_nchw_wide(cuda):
    _run(cuda, _lib.DECONV_5X5_S2, 1, 128, 192, 8, 8, out_nchw=True)


def test_persistent_many_tiles(cuda):
    # more tiles than SMs: every CTA loops, ring phases wrap many times
    _run(cuda, _lib.CONV_5X5_S2, 24, 128, 128, 64, 64, epi=_lib.EPI_GDN)


def test_unsupported_shapes_fail_loudly(cuda):
    x = torch.zeros(1, 4, 4, 100, dtype=torch.bfloat16, device=cuda)
    w = torch.zeros(25 * 128 * 128, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(NotImplementedError):
        ops.conv_forward(x, kind=_lib.CONV_5X5_S2, epilogue=0, in_layout=1, out_layout=1, in_c=100, out_c=128,
                         weight=w, bias=None)
    with pytest.raises(RuntimeError):
        ops.nchw_to_nhwc_bf16(torch.zeros(1, 3, 4, 4))


def test_first_layer_pipelined_variants(cuda):
    # conv_first2_kernel: both band counts, both widths, every epilogue, partial tiles (OH, OW not multiples of 8 / 16),
    # more tiles than SMs (ring slots and TMEM buffers wrap many times)
    _run(cuda, _lib.CONV_5X5_S2, 1, 1, 64, 32, 36, epi=_lib.EPI_GDN, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 2, 3, 64, 40, 24, epi=_lib.EPI_RELU, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 1, 3, 128, 40, 24, epi=_lib.EPI_IGDN, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 1, 1, 128, 24, 52, epi=_lib.EPI_NONE, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 6, 3, 128, 128, 128, epi=_lib.EPI_GDN, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 5, 1, 128, 128, 160, epi=_lib.EPI_GDN, first=True)
    # N = 192 (hyperprior widths): one epilogue team, two TMEM accumulators
    _run(cuda, _lib.CONV_5X5_S2, 2, 3, 192, 64, 72, epi=_lib.EPI_GDN, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 1, 1, 192, 40, 32, epi=_lib.EPI_RELU, first=True)
    _run(cuda, _lib.CONV_5X5_S2, 3, 3, 192, 160, 128, epi=_lib.EPI_GDN, first=True)


def test_last_layer_segments_strips_and_widths(cuda):
    # deconv_narrow2_kernel: two 128-pixel row segments, a ragged last strip (H = 32 + 1), 3- and 4-chunk inputs,
    # more strips than SMs (TMEM slot ring wraps), out_c = 2 and 4
    _run(cuda, _lib.DECONV_5X5_S2, 1, 128, 3, 40, 200, out_nchw=True)
    _run(cuda, _lib.DECONV_5X5_S2, 1, 256, 2, 33, 20, out_nchw=True)
    _run(cuda, _lib.DECONV_5X5_S2, 2, 192, 4, 9, 130, epi=_lib.EPI_RELU, out_nchw=True)
    _run(cuda, _lib.DECONV_5X5_S2, 40, 64, 1, 128, 16, out_nchw=True)


def test_kernels_are_safe_under_concurrent_streams(cuda):
    # bench.py's host-buffer leg runs three streams at once: results must not depend on what else is on the GPU
    g = torch.Generator().manual_seed(3)
    jobs = []
    for kind, cin, cout, H, W, epi, first, nchw in [
            (_lib.CONV_5X5_S2, 3, 128, 64, 64, _lib.EPI_GDN, True, False),
            (_lib.CONV_5X5_S2, 128, 128, 32, 32, _lib.EPI_GDN, False, False),
            (_lib.DECONV_5X5_S2, 128, 128, 16, 16, _lib.EPI_IGDN, False, False),
            (_lib.DECONV_5X5_S2, 128, 3, 32, 32, _lib.EPI_NONE, False, True)]:
        x = torch.randn(24, cin, H, W, generator=g)
        wshape = (cin, cout, 5, 5) if kind == _lib.DECONV_5X5_S2 else (cout, cin, 5, 5)
        w = torch.randn(wshape, generator=g) / (cin * 25) ** 0.5
        in_layout = _lib.LAYOUT_NCHW_F32 if first else _lib.LAYOUT_NHWC_BF16
        out_layout = _lib.LAYOUT_NCHW_F32 if nchw else _lib.LAYOUT_NHWC_BF16
        xd = x.to(cuda) if first else x.permute(0, 2, 3, 1).contiguous().bfloat16().to(cuda)
        packed = ops.pack_conv_weight(w.to(cuda), kind, cout, cin, in_layout)
        bh = gh = None
        if epi in (_lib.EPI_GDN, _lib.EPI_IGDN):
            bh, gh = ops.gdn_pack(torch.ones(cout, device=cuda), (0.1 * torch.eye(cout)).sqrt().to(cuda), 0.0, 0.0, 0.0)
        kw = dict(kind=kind, epilogue=epi, in_layout=in_layout, out_layout=out_layout, in_c=cin, out_c=cout,
                  weight=packed, bias=torch.zeros(cout, device=cuda), beta=bh, gamma=gh)
        jobs.append((xd, kw, ops.conv_forward(xd, **kw).float().clone()))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda) for _ in range(3)]
    for it in range(30):
        outs = []
        for s in streams:
            with torch.cuda.stream(s):
                for xd, kw, ref in jobs:
                    outs.append((ops.conv_forward(xd, **kw), ref))
        torch.cuda.synchronize()
        for o, ref in outs:
            assert torch.equal(o.float(), ref), "a kernel's result changed under concurrent load"
