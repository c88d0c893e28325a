"""-m gpu: the memory-bound kernels (quantise, likelihoods, symbols, CDF indexes, reductions, layout) through
the C ABI against the CPU oracle and the committed golden vectors.  Integer outputs: bit-exact.  Likelihoods:
max-abs <= 1e-6 (fp32; the only differences are 1-ulp expf/tanhf/erfcf library differences)."""
import math
import os

import numpy as np
import pytest
import torch

import licos_b200 as L
from licos_b200 import ops, synth
from oracle import compressai_ref as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LIK_TOL = 1e-6


def _pair(in_ch, device):
    torch.manual_seed(42)  # same construction order as tests/golden/make_golden.py (no surgery for RGB)
    if in_ch == 3:
        net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
        ref = R.image_models["bmshj2018-factorized"](quality=1)
    else:
        net = L.get_model("bmshj2018-factorized", False, in_ch, 1)
        ref = R.get_model("bmshj2018-factorized", False, in_ch, 1)
    synth.condition_weights(net)
    ref.load_state_dict(net.state_dict())
    return net.to(device).eval(), ref.eval()


@pytest.mark.parametrize("tag,in_ch", [("rgb", 3), ("split", 1), ("merged", 13)])
def test_eb_forward_matches_golden(cuda, tag, in_ch):
    z = np.load(os.path.join(GOLD, f"factorized_{tag}.npz"))
    net, _ = _pair(in_ch, cuda)
    eb = net.entropy_bottleneck
    y = torch.from_numpy(z["y"]).to(cuda)
    with torch.no_grad():
        y_hat, lik = eb(y)
        sym = eb.symbols(y)
        y_noisy, lik_noisy = eb(y, training=True, noise=torch.from_numpy(z["noise"]).to(cuda))
    assert np.array_equal(sym.cpu().numpy(), z["symbols"])
    assert np.array_equal(y_hat.cpu().numpy(), z["y_hat"])
    assert np.abs(lik.cpu().numpy() - z["lik"]).max() <= LIK_TOL
    assert np.array_equal(y_noisy.cpu().numpy(), z["y_noisy"])
    assert np.abs(lik_noisy.cpu().numpy() - z["lik_noisy"]).max() <= LIK_TOL


@pytest.mark.parametrize("form", ["plain", "stable"])
def test_eb_ties_escapes_and_ragged_shapes(cuda, form):
    net, ref = _pair(3, cuda)
    eb, reb = net.entropy_bottleneck, ref.entropy_bottleneck
    eb.likelihood_form = reb.likelihood_form = form
    eb._packed_key = None
    med = reb.quantiles[:, 0, 1].detach()
    # ties y - med in {+-0.5, +-1.5, +-2.5}, values far outside the LUT / CDF support, an odd spatial size
    base = torch.tensor([0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.0, 127.5, 128.5, -128.5, 129.0, 300.0, -777.25, 1e4])
    g = torch.Generator().manual_seed(0)
    y = torch.randn(3, 192, 5, 3, generator=g) * 6
    y.view(3, 192, -1)[:, :, : base.numel()] = base[None, None, :] + med[None, :, None]
    for yy in (y, y[:1, :, :1, :1].contiguous(), y[:0]):
        with torch.no_grad():
            a_hat, a_lik = eb(yy.to(cuda))
            r_hat, r_lik = reb(yy)
            sym, idx = ops.eb_symbols(yy.to(cuda), eb.packed_params().medians, want_indexes=True)
            r_sym = reb.quantize(yy, "symbols", med.reshape(1, -1, 1, 1))
        assert torch.equal(a_hat.cpu(), r_hat)
        assert torch.equal(sym.cpu(), r_sym)
        assert torch.equal(idx.cpu(), reb._build_indexes(yy.size()))
        if yy.numel():
            assert (a_lik.cpu() - r_lik).abs().max() <= LIK_TOL
            assert torch.equal(ops.eb_dequantize(sym, eb.packed_params().medians).cpu(),
                               reb.dequantize(r_sym, med.reshape(1, -1, 1, 1).expand_as(yy)))
    # round-half-even at the quantiser
    t = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5]).reshape(1, 1, 5, 1).repeat(1, 192, 1, 1).to(cuda)
    s = ops.eb_symbols(t, torch.zeros(192, device=cuda))
    assert s[0, 0, :, 0].tolist() == [0, 2, 2, 0, -2]


def test_eb_noise_philox_statistics(cuda):
    net, _ = _pair(3, cuda)
    eb = net.entropy_bottleneck
    y = torch.zeros(4, 192, 32, 32, device=cuda)
    with torch.no_grad():
        a, lik = eb(y, training=True, seed=123)
        b, _ = eb(y, training=True, seed=123)
        c, _ = eb(y, training=True, seed=124)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert a.min() >= -0.5 and a.max() < 0.5
    assert abs(a.mean().item()) < 2e-3 and abs(a.var().item() - 1 / 12) < 2e-3
    assert (lik > 0).all()


def test_gaussian_conditional_matches_golden_and_oracle(cuda):
    z = np.load(os.path.join(GOLD, "hyperprior_rgb.npz"))
    gc = L.GaussianConditional(None)
    gc.update_scale_table(L.get_scale_table())
    gc = gc.to(cuda).eval()
    assert np.array_equal(gc._quantized_cdf[0, :5].cpu().numpy(), z["gc_cdf_row0"])
    y, s = torch.from_numpy(z["y"]).to(cuda), torch.from_numpy(z["scales"]).to(cuda)
    with torch.no_grad():
        y_hat, lik = gc(y, s)
        idx = gc.build_indexes(s)
    assert np.array_equal(idx.cpu().numpy(), z["indexes"])
    assert np.array_equal(y_hat.cpu().numpy(), z["y_hat"])
    assert np.abs(lik.cpu().numpy() - z["y_lik"]).max() <= LIK_TOL
    # sigma exactly on table entries, both clamps, means, given noise
    rgc = R.GaussianConditional(None).eval()
    rgc.update_scale_table(R.get_scale_table())
    t = rgc.scale_table
    s2 = torch.cat([t, t * (1 + 1e-6), t * (1 - 1e-6), torch.tensor([0.0, 0.05, 0.11, 1e3, 1e6])]).reshape(1, 1, -1, 1)
    g = torch.Generator().manual_seed(1)
    y2 = torch.randn(s2.shape, generator=g) * 20
    mu = torch.randn(s2.shape, generator=g)
    nz = torch.rand(s2.shape, generator=g) - 0.5
    with torch.no_grad():
        assert torch.equal(gc.build_indexes(s2.to(cuda)).cpu(), rgc.build_indexes(s2))
        for kw, rkw in (({}, {}), ({"means": mu.to(cuda)}, {"means": mu}),
                        ({"training": True, "noise": nz.to(cuda)}, {"training": True, "noise": nz})):
            a_hat, a_lik = gc(y2.to(cuda), s2.to(cuda), **kw)
            r_hat, r_lik = rgc(y2, s2, **rkw)
            assert torch.equal(a_hat.cpu(), r_hat)
            assert (a_lik.cpu() - r_lik).abs().max() <= LIK_TOL
        assert torch.equal(gc.quantize(y2.to(cuda), "symbols", mu.to(cuda)).cpu(), rgc.quantize(y2, "symbols", mu))
        assert torch.equal(gc.quantize(y2.to(cuda), "symbols").cpu(), rgc.quantize(y2, "symbols"))


def test_reductions_and_rd_loss(cuda):
    g = torch.Generator().manual_seed(3)
    lik = (torch.rand(3, 192, 16, 16, generator=g) * 0.9 + 1e-9)
    a, b = torch.rand(3, 3, 256, 256, generator=g), torch.rand(3, 3, 256, 256, generator=g)
    out = {"x_hat": a.to(cuda), "likelihoods": {"y": lik.to(cuda), "z": lik[:, :7].contiguous().to(cuda)}}
    rout = {"x_hat": a, "likelihoods": {"y": lik, "z": lik[:, :7]}}
    with torch.no_grad():
        got = L.RateDistortionLoss(lmbda=1e-2)(out, b.to(cuda))
    exp = R.RateDistortionLoss(lmbda=1e-2)(rout, b)
    for k in ("bpp_loss", "mse_loss", "loss"):
        assert abs(got[k].item() - exp[k].item()) <= 2e-6 * abs(exp[k].item()), k
    assert abs(L.compute_bpp(out) - exp["bpp_loss"].item()) < 1e-5
    assert abs(L.compute_psnr(out["x_hat"], b.to(cuda)) - (-10 * math.log10(exp["mse_loss"].item()))) < 1e-4


def test_layout_roundtrip(cuda):
    g = torch.Generator().manual_seed(4)
    for shape in ((2, 192, 16, 16), (1, 64, 5, 7), (3, 320, 1, 1)):
        x = torch.randn(shape, generator=g)
        nhwc = ops.nchw_to_nhwc_bf16(x.to(cuda))
        assert torch.equal(nhwc.cpu(), x.permute(0, 2, 3, 1).contiguous().bfloat16())
        assert torch.equal(ops.nchw_to_nhwc_bf16(x.to(cuda), take_abs=True).cpu(),
                           x.abs().permute(0, 2, 3, 1).contiguous().bfloat16())
        assert torch.equal(ops.nhwc_bf16_to_nchw(nhwc).cpu(), x.bfloat16().float())


def test_federated_pair_merge_on_device(cuda):
    torch.manual_seed(5)
    a = L.get_model("bmshj2018-factorized", False, 1, 1).to(cuda)
    b = L.get_model("bmshj2018-factorized", False, 1, 1).to(cuda)
    from licos_b200.federated import merge_pair
    got = merge_pair(a.state_dict(), b.state_dict(), loss=3.0, best_loss=1.0)
    cpu = lambda sd: {k: v.cpu() for k, v in sd.items()}
    exp = R.federated_average(cpu(a.state_dict()), cpu(b.state_dict()), loss=3.0, best_loss=1.0)
    for k, v in exp.items():
        assert torch.equal(got[k].cpu(), v), k  # same two roundings as `w*a; += w*b`


@pytest.mark.parametrize("tag,in_ch", [("rgb", 3), ("split", 1)])
def test_eb_forward_fused_equals_separate_kernels_and_golden(cuda, tag, in_ch):
    # one pass over y: y_hat / likelihoods / symbols bit-identical to the separate kernels (and to the golden vectors),
    # the bf16 NHWC copy equal to the rounded, transposed y_hat; ties, escapes, a partial 64-position tile, odd channels
    z = np.load(os.path.join(GOLD, f"factorized_{tag}.npz"))
    net, _ = _pair(in_ch, cuda)
    eb = net.entropy_bottleneck
    y = torch.from_numpy(z["y"]).to(cuda)
    with torch.no_grad():
        y_hat, lik, sym, nhwc = eb.forward_fused(y, want_symbols=True, want_nhwc=True)
        y_hat0, lik0 = eb(y)
    assert np.array_equal(sym.cpu().numpy(), z["symbols"])
    assert np.array_equal(y_hat.cpu().numpy(), z["y_hat"])
    assert torch.equal(lik, lik0) and torch.equal(y_hat, y_hat0)
    assert torch.equal(nhwc, y_hat.permute(0, 2, 3, 1).contiguous().bfloat16())

    med = eb.quantiles[:, 0, 1].detach().reshape(1, -1, 1, 1)
    g = torch.Generator().manual_seed(1)
    for shape in [(3, 192, 12, 20), (2, 192, 2, 2), (1, 192, 5, 3)]:   # 240 = 3 tiles + 48; 4; 15 (falls back)
        yy = (torch.randn(shape, generator=g) * 40).to(cuda)
        yy[0, :, 0, 0] = med[0, :, 0, 0] + 0.5           # ties
        yy[0, :, 0, 1] = med[0, :, 0, 0] - 1.5
        yy[-1, :, -1, -1] = 500.0                        # beyond the likelihood table
        with torch.no_grad():
            a = eb.forward_fused(yy, want_symbols=True, want_nhwc=True)
            b = eb(yy)
            s = eb.symbols(yy)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], s)
        assert torch.equal(a[3], b[0].permute(0, 2, 3, 1).contiguous().bfloat16())
    with torch.no_grad():
        out = net(synth.make_input("rgb256" if in_ch == 3 else "raw512", 2).to(cuda))   # model.forward uses the fused pass
    assert out["x_hat"].shape[0] == 2


def test_device_rans_encoder_is_bitstream_identical(cuda):
    # the GPU encoder against the host coder (itself pinned to the C / Python oracle twins): same bytes for ordinary
    # symbols, out-of-support symbols (bypass nibbles, long base-15 prefixes), a shared and a per-image index plane,
    # odd lengths (scalar tail of the 4-wide backwards walk) and a stream that does not fit its scratch row
    net, ref = _pair(3, cuda)
    net.update()
    eb = net.entropy_bottleneck
    cdf, lens, offs = eb._quantized_cdf, eb._cdf_length, eb._offset
    g = torch.Generator().manual_seed(7)
    for shape, scale in [((5, 192, 16, 16), 6.0), ((3, 192, 3, 5), 30.0), ((33, 192, 4, 4), 2.0), ((2, 192, 1, 1), 2000.0)]:
        y = (torch.randn(shape, generator=g) * scale).to(cuda)
        y[0, 0, 0, 0] = 3.0e5                      # 5 nibbles of bypass
        y[-1, -1, -1, -1] = -1.0e6                 # 6 nibbles
        sym = eb.symbols(y)
        B, n_sp = shape[0], shape[2] * shape[3]
        host = ops.rans_encode_batch(sym.reshape(B, -1).cpu().numpy(),
                                     np.repeat(np.arange(192, dtype=np.int32), n_sp), cdf.cpu().numpy(), lens.cpu().numpy(),
                                     offs.cpu().numpy())
        dev_implicit = ops.rans_encode_device(sym, None, n_sp, cdf, lens, offs)
        idx_plane = torch.arange(192, dtype=torch.int32, device=cuda).repeat_interleave(n_sp)
        dev_shared = ops.rans_encode_device(sym, idx_plane, n_sp, cdf, lens, offs)
        dev_full = ops.rans_encode_device(sym, idx_plane.repeat(B, 1), n_sp, cdf, lens, offs)
        assert dev_implicit == host and dev_shared == host and dev_full == host
        eb.device_coder = False
        via_host = eb.compress(y)
        eb.device_coder = True
        assert eb.compress(y) == via_host == host
        assert torch.equal(eb.decompress(host, shape[2:]), eb(y)[0])
        # the GPU decoder: same symbols back, through every index form; and through the host decoder
        flat = sym.reshape(B, -1)
        for ix in (None, idx_plane, idx_plane.repeat(B, 1)):
            dec = ops.rans_decode_device(host, ix, flat.shape[1], n_sp, cdf, lens, offs)
            assert dec is not None and torch.equal(dec, flat)
        eb.device_coder = False
        y_host = eb.decompress(host, shape[2:])
        eb.device_coder = True
        assert torch.equal(eb.decompress(host, shape[2:]), y_host)
    assert ops.rans_decode_device([b"\x00" * 6], None, 4, 1, cdf, lens, offs) is None   # not a whole number of words


def test_ms_ssim_on_device_matches_oracle(cuda):
    # eval_utils.py:159-169 compute_msssim: device kernels vs the CPU restatement of pytorch_msssim.ms_ssim; odd sizes
    # exercise the padded average pooling; identical images give exactly 1
    from oracle import msssim_ref as M

    g = torch.Generator().manual_seed(0)
    for shape, noise in [((2, 3, 256, 256), 0.05), ((1, 1, 177, 231), 0.2), ((3, 2, 161, 170), 0.01)]:
        x = torch.rand(shape, generator=g)
        y = (x + noise * torch.randn(shape, generator=g)).clamp(0, 1)
        want = M.ms_ssim(x, y, data_range=1.0).item()
        got = ops.ms_ssim(x.to(cuda), y.to(cuda), data_range=1.0).item()
        assert abs(got - want) <= 2e-5, (shape, got, want)
        want255 = M.ms_ssim(x * 255, y * 255, data_range=255.0).item()
        got255 = ops.ms_ssim((x * 255).to(cuda), (y * 255).to(cuda), data_range=255.0).item()
        assert abs(got255 - want255) <= 2e-5
    assert abs(L.compute_msssim(x.to(cuda), x.to(cuda)) - 1.0) <= 1e-6
    with pytest.raises(AssertionError):
        ops.ms_ssim(torch.rand(1, 1, 160, 300, device=cuda), torch.rand(1, 1, 160, 300, device=cuda))


def test_raw_dn_scaling_matches_reference_arithmetic(cuda):
    # raw_image_folder.py:192-196: band = dn / 4095 in float64; img_as_ubyte(band) / 255 unless use_full_range; float32
    dn = np.arange(0, 4096, dtype=np.uint16)
    dn = np.concatenate([dn, np.random.default_rng(0).integers(0, 4096, 10000).astype(np.uint16)])
    band = dn / 4095
    want_full = band.astype(np.float32)
    want_8bit = (np.clip(np.rint(band * 255), 0, 255).astype(np.uint8) / 255).astype(np.float32)
    t = torch.from_numpy(dn.view(np.int16)).to(cuda)
    assert np.array_equal(ops.raw_dn_to_unit(t, use_full_range=True).cpu().numpy(), want_full)
    assert np.array_equal(ops.raw_dn_to_unit(t, use_full_range=False).cpu().numpy(), want_8bit)
