"""-m gpu: integer tiles straight into the first layer and out of the last one (SURVEY.md section 8f row N4,
/root/reference/licos/raw_image_folder.py:192-196: the reference scales DN / DN_MAX on the host and ships fp32).

The bar is BIT-EQUALITY with this package's own fp32 path fed the reference's host-side scaling -- the integer layouts
change what crosses PCIe, never a result:
  * first layer:  conv(u8 / u16 / u16+8-bit requant tiles)  ==  conv(float32(float64(v) / int_max))
  * last layer:   u8 / u16 x_hat                            ==  round(clamp(x_hat_fp32, 0, 1) * int_max)
  * bottleneck:   int16 symbols == int32 symbols;  fused rate term == sum(log(likelihoods))
  * end to end:   forward_tiles(u8) == forward(fp32 of the same pixels), symbols and images identical
"""
import math

import numpy as np
import pytest
import torch

import licos_b200 as L
from licos_b200 import _lib, ops, synth

pytestmark = pytest.mark.gpu


def _unit(v: torch.Tensor, int_max: int, requant8: bool = False) -> torch.Tensor:
    """raw_image_folder.py:192-196 on the host: float64 division, optional img_as_ubyte step, float32 tensor."""
    a = v.cpu().numpy().astype(np.float64) / int_max
    if requant8:
        a = np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8) / 255
    return torch.from_numpy(a.astype(np.float32))


def _as_int(t: torch.Tensor) -> torch.Tensor:
    """int32 values of a uint8 / uint16 tensor (uint16 has few torch kernels: go through its int16 view)."""
    if t.dtype == torch.uint16:
        return t.view(torch.int16).to(torch.int32) & 0xFFFF
    return t.to(torch.int32)


def _first_layer(cuda, cin, cout, epi, x, **kw):
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(cin, cout, 5, 2, 2).to(cuda)
    gdn = L.GDN(cout).to(cuda)
    seq = L.FusedSequential(conv, gdn) if epi == "gdn" else L.FusedSequential(conv, torch.nn.ReLU())
    # a second layer so that the first one writes bf16 NHWC (the pipelined kernel's output layout)
    seq = L.FusedSequential(*seq, torch.nn.Conv2d(cout, 64, 5, 2, 2).to(cuda))
    with torch.no_grad():
        return seq(x.to(cuda), **kw)


@pytest.mark.parametrize("cin,cout,epi,hw", [(3, 128, "gdn", (64, 96)), (1, 128, "gdn", (72, 48)), (3, 192, "gdn", (48, 64)),
                                             (1, 192, "relu", (40, 80)), (3, 128, "relu", (34, 112)), (1, 64, "gdn", (130, 16))])
@pytest.mark.parametrize("mode", ["u8", "u16", "u16_q8", "u16_1023"])
def test_first_layer_integer_tiles_bit_equal_to_fp32_path(cuda, cin, cout, epi, hw, mode):
    g = torch.Generator().manual_seed(11)
    if mode == "u8":
        v, int_max, kw = torch.randint(0, 256, (2, cin, *hw), generator=g).to(torch.uint8), 255, {}
    elif mode == "u16_1023":
        v, int_max, kw = torch.randint(0, 1024, (2, cin, *hw), generator=g).to(torch.int16), 1023, {"int_max": 1023}
    else:
        v, int_max = torch.randint(0, 4096, (2, cin, *hw), generator=g).to(torch.int16), 4095
        v[0, 0, 0, :4] = torch.tensor([0, 4095, 2047, 2048], dtype=torch.int16)
        kw = {"requant8": True} if mode == "u16_q8" else {}
    want = _first_layer(cuda, cin, cout, epi, _unit(v, int_max, mode == "u16_q8"))
    got = _first_layer(cuda, cin, cout, epi, v, **kw)
    assert torch.equal(got, want)
    if mode != "u8":  # uint16 storage of the same numbers
        got16 = _first_layer(cuda, cin, cout, epi, v.view(torch.uint16), **kw)
        assert torch.equal(got16, want)


def test_integer_tiles_other_shapes_take_the_standalone_scaling(cuda):
    """13 bands (the reference's "merged" raw format) and rows that are not 16-byte multiples: same values via
    licos_raw_dn_to_unit + the fp32 path."""
    g = torch.Generator().manual_seed(12)
    for cin, hw in ((13, (32, 48)), (3, (32, 36))):
        v = torch.randint(0, 4096, (1, cin, *hw), generator=g).to(torch.int16)
        torch.manual_seed(4)
        net = L.get_model("bmshj2018-factorized", False, cin, 1).to(cuda).eval()
        with torch.no_grad():
            assert torch.equal(net.g_a(v.to(cuda)), net.g_a(_unit(v, 4095).to(cuda)))
            assert torch.equal(net.g_a(v.to(cuda), requant8=True), net.g_a(_unit(v, 4095, True).to(cuda)))


def test_refused_full_scale_values(cuda):
    assert _lib.lib.licos_pixel_scale_exact(4095, 0) == 1 and _lib.lib.licos_pixel_scale_exact(4095, 1) == 1
    assert _lib.lib.licos_pixel_scale_exact(4096, 1) == 0  # even full scale: exact ties in v / max * 255
    v = torch.zeros(1, 3, 32, 32, dtype=torch.int16, device=cuda)
    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False).to(cuda).eval()
    with pytest.raises(NotImplementedError):
        net.g_a(v, int_max=4096, requant8=True)
    with pytest.raises(NotImplementedError):  # training on integer tiles is not a path
        net.g_a.train()
        net.g_a(v)


@pytest.mark.parametrize("cout,dtype,out_max,hw", [(3, torch.uint8, 255, (24, 40)), (1, torch.uint16, 4095, (17, 130)),
                                                   (3, torch.uint16, 65535, (8, 8)), (4, torch.uint8, 100, (33, 20))])
def test_last_layer_integer_pixels_equal_rounded_fp32(cuda, cout, dtype, out_max, hw):
    torch.manual_seed(5)
    seq = L.FusedSequential(torch.nn.ConvTranspose2d(128, cout, 5, 2, 2, 1)).to(cuda)
    with torch.no_grad():
        seq[0].weight.mul_(6.0)  # both clamps are hit
        seq[0].bias.add_(0.5)
    g = torch.Generator().manual_seed(13)
    y = torch.randn(2, 128, *hw, generator=g).to(cuda)
    with torch.no_grad():
        f = seq(y)
        q = seq(y, out_dtype=dtype, out_max=out_max)
    assert q.dtype == dtype and q.shape == f.shape
    want = torch.round(f.clamp(0, 1) * float(out_max))
    assert (want == 0).any() and (want == out_max).any()
    assert torch.equal(_as_int(q).float(), want)


def test_fused_bottleneck_int16_symbols_rate_term_and_cached_table(cuda):
    torch.manual_seed(42)
    net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
    synth.condition_weights(net)
    eb = net.entropy_bottleneck.to(cuda).eval()
    g = torch.Generator().manual_seed(14)
    y = (torch.randn(3, 192, 8, 12, generator=g) * 6).to(cuda)
    y[0, 0, 0, :3] = torch.tensor([40000.0, -40000.0, 300.0])  # saturation of the int16 copy, the direct likelihood path
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    with torch.no_grad():
        y_hat, lik, sym, nhwc = eb.forward_fused(y, want_symbols=True, want_nhwc=True)
        y_hat2, lik2, sym2, nhwc2, sym16 = eb.forward_fused(y, want_symbols=True, want_nhwc=True, want_symbols_i16=True, sum_ln=acc)
        none_a, none_b, _, nhwc3, sym16b = eb.forward_fused(y, want_nhwc=True, want_symbols_i16=True, want_float=False)
        ref_hat, ref_lik = eb(y)
    assert torch.equal(y_hat, ref_hat) and torch.equal(lik, ref_lik)
    assert torch.equal(y_hat2, y_hat) and torch.equal(lik2, lik) and torch.equal(sym2, sym) and torch.equal(nhwc2, nhwc)
    assert none_a is None and none_b is None and torch.equal(nhwc3, nhwc) and torch.equal(sym16b, sym16)
    assert torch.equal(sym16.to(torch.int32), sym.clamp(-32768, 32767))
    assert int(sym16[0, 0, 0, 0]) == 32767 and int(sym16[0, 0, 0, 1]) == -32768
    want = torch.log(lik.double()).sum().item()
    assert abs(acc.item() - want) <= 1e-6 * abs(want)
    # the cached table follows the parameters
    with torch.no_grad():
        eb._bias0.add_(0.05)
        _, lik_new, _, _ = eb.forward_fused(y)
        assert not torch.equal(lik_new, lik) and torch.equal(lik_new, eb(y)[1])


@pytest.mark.parametrize("cin,kind", [(3, "u8"), (1, "u16")])
def test_forward_tiles_equals_forward_on_the_same_pixels(cuda, cin, kind):
    torch.manual_seed(42)
    net = L.get_model("bmshj2018-factorized", False, cin, 1)
    synth.condition_weights(net)
    net = net.to(cuda).eval()
    g = torch.Generator().manual_seed(15)
    if kind == "u8":
        v, int_max, out_dtype = torch.randint(0, 256, (4, cin, 128, 64), generator=g).to(torch.uint8), 255, torch.uint8
    else:
        v, int_max, out_dtype = torch.randint(0, 4096, (4, cin, 64, 128), generator=g).to(torch.int16), 4095, torch.uint16
    x = _unit(v, int_max).to(cuda)
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    with torch.no_grad():
        ref = net(x)
        sym = net.entropy_bottleneck.symbols(net.g_a(x))
        out = net.forward_tiles(v.to(cuda), out_dtype=out_dtype, sum_ln=acc)
        outf = net.forward_tiles(v.to(cuda))
        lean = net.forward_tiles(v.to(cuda), out_dtype=out_dtype, want_float=False)
    assert torch.equal(outf["x_hat"], ref["x_hat"]) and torch.equal(outf["likelihoods"]["y"], ref["likelihoods"]["y"])
    assert torch.equal(out["symbols"].to(torch.int32), sym)
    assert out["x_hat"].dtype == out_dtype
    assert torch.equal(_as_int(out["x_hat"]).float(), torch.round(ref["x_hat"].clamp(0, 1) * float(int_max)))
    assert torch.equal(lean["x_hat"], out["x_hat"]) and torch.equal(lean["symbols"], out["symbols"])
    assert lean["likelihoods"]["y"] is None
    n_px = v.shape[0] * v.shape[2] * v.shape[3]
    assert abs(acc.item() / (-math.log(2) * n_px) - L.compute_bpp(ref)) <= 1e-6
