"""-m gpu: the native backward pass (licos/train.py:193 `loss.backward()`) against CPU fp32/fp64 autograd.

Kernel level: licos_conv_wgrad and the GDN / ReLU / bias pieces against torch expressions evaluated on the SAME
bf16-rounded operands (so the only difference is the fp32 accumulation order): relative error <= 2e-3 of the result's scale.
Chain level: g_a / g_s / h_a / h_s parameter and input gradients against the oracle's fp32 autograd with the same
state_dict.  Activations and their gradients are bf16 on the device, so the bar is the one bf16 training has:
cosine similarity >= 0.999 and relative L2 error <= 3e-2 per tensor (both printed).  Chains with ReLU get 0.995 / 0.1 (0.99 / 0.15 for the four-ReLU g_a):
a pre-activation within bf16 rounding of zero opens or closes its gate differently from the fp32 oracle."""
import pytest
import torch
import torch.nn.functional as F

import licos_b200 as L
from licos_b200 import _lib, ops, synth
from oracle import compressai_ref as R

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


def _rel(got, ref):
    return (got.double().cpu() - ref.double()).abs().max().item() / max(ref.double().abs().max().item(), 1e-30)


def _wgrad_ref(small, big, kind):
    """small (B,h,w,cs), big (B,H,W,cb) bf16-rounded values as float64 CPU -> [taps][cs][cb]."""
    s = small.permute(0, 3, 1, 2).double()
    b = big.permute(0, 3, 1, 2).double()
    cs, cb = s.shape[1], b.shape[1]
    if kind in (_lib.CONV_5X5_S2, _lib.DECONV_5X5_S2):
        w = torch.nn.grad.conv2d_weight(b, (cs, cb, 5, 5), s, stride=2, padding=2)
    elif kind == _lib.CONV_3X3_S1:
        w = torch.nn.grad.conv2d_weight(b, (cs, cb, 3, 3), s, stride=1, padding=1)
    else:
        w = torch.einsum("bshw,bchw->sc", s, b).reshape(cs, cb, 1, 1)
    k = w.shape[2]
    return w.permute(2, 3, 0, 1).reshape(k * k, cs, cb)


@pytest.mark.parametrize("kind,cs,cb,hw,big_hw,batch", [
    (_lib.CONV_5X5_S2, 128, 128, (16, 24), (32, 48), 2),
    (_lib.CONV_5X5_S2, 192, 128, (8, 16), (16, 32), 3),
    (_lib.DECONV_5X5_S2, 128, 192, (12, 20), (24, 40), 2),
    (_lib.CONV_5X5_S2, 128, 128, (9, 5), (17, 10), 2),
    (_lib.CONV_3X3_S1, 128, 320, (16, 16), (16, 16), 2),
    (_lib.CONV_1X1, 128, 128, (32, 32), (32, 32), 2),
    (_lib.CONV_1X1, 192, 64, (10, 7), (10, 7), 5),
    (_lib.CONV_1X1, 128, 80, (24, 20), (24, 20), 3),     # 3-band patch matrix: 75 columns padded to 80
    (_lib.CONV_1X1, 128, 32, (9, 33), (9, 33), 2),       # 1 band: 25 -> 32
    (_lib.CONV_1X1, 128, 336, (8, 16), (8, 16), 2),      # 13 bands: 325 -> 336 = 128 + 128 + 80
    (_lib.CONV_5X5_S2, 128, 128, (64, 64), (128, 128), 8),
])
def test_wgrad_vs_cpu(cuda, kind, cs, cb, hw, big_hw, batch):
    g = torch.Generator().manual_seed(hw[0] * 31 + cs)
    small = _bf(torch.randn(batch, *hw, cs, generator=g))
    big = _bf(torch.randn(batch, *big_hw, cb, generator=g))
    got = ops.conv_wgrad(small.to(cuda).contiguous(), big.to(cuda).contiguous(), kind)
    torch.cuda.synchronize()
    ref = _wgrad_ref(small.float(), big.float(), kind)
    err = _rel(got, ref)
    print(f"wgrad kind {kind} cs {cs} cb {cb} {hw}: max|err|/max = {err:.3e}")
    assert got.shape == ref.shape
    assert err <= 2e-3


def test_wgrad_accumulates_and_repeats(cuda):
    g = torch.Generator().manual_seed(3)
    small = _bf(torch.randn(4, 32, 32, 128, generator=g)).to(cuda)
    big = _bf(torch.randn(4, 64, 64, 128, generator=g)).to(cuda)
    a = ops.conv_wgrad(small, big, _lib.CONV_5X5_S2)
    b = ops.conv_wgrad(small, big, _lib.CONV_5X5_S2)
    # red.add order differs from run to run: equal to fp32 rounding of the partial sums, not bit-equal
    assert _rel(a, b.cpu()) <= 1e-5


@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_backward_pieces(cuda, inverse):
    g = torch.Generator().manual_seed(5)
    n = 4 * 16 * 16 * 128
    x = _bf(torch.randn(n, generator=g))
    gr = _bf(torch.randn(n, generator=g))
    norm = _bf(torch.rand(n, generator=g) * 3 + 0.5)
    t = _bf(torch.randn(n, generator=g))
    x2 = ops.square_bf16(x.to(cuda))
    assert torch.equal(x2.cpu(), _bf(x.float() ** 2))
    xs = x.to(cuda).reshape(4, 16, 16, 128)
    sum_dn = torch.zeros(128, device=cuda)
    d_norm, d_direct = ops.gdn_bwd_mid(xs, gr.to(cuda).reshape(xs.shape), norm.to(cuda).reshape(xs.shape), inverse, sum_out=sum_dn)
    assert _rel(sum_dn, d_norm.float().reshape(-1, 128).double().sum(0).cpu()) <= 1e-4
    d_norm, d_direct = d_norm.reshape(-1), d_direct.reshape(-1)
    xf, gf, nf = x.float(), gr.float(), norm.float()
    if inverse:
        ref_dd, ref_dn = gf * nf.sqrt(), 0.5 * gf * xf / nf.sqrt()
    else:
        ref_dd, ref_dn = gf * nf.rsqrt(), -0.5 * gf * xf * nf.rsqrt() ** 3
    assert _rel(d_direct.float(), _bf(ref_dd).float()) <= 1e-2
    assert _rel(d_norm.float(), _bf(ref_dn).float()) <= 1e-2
    dd = d_direct.clone()
    sum_dx = torch.zeros(128, device=cuda)
    dx = ops.gdn_bwd_out(xs, t.to(cuda).reshape(xs.shape), dd.reshape(xs.shape), sum_out=sum_dx).reshape(-1)
    assert _rel(sum_dx, dx.float().reshape(-1, 128).double().sum(0).cpu()) <= 1e-4
    x192 = x[:192 * 600].reshape(1, 20, 30, 192).to(cuda)   # a thread count that is not a multiple of C / 8 by default
    s192 = torch.zeros(192, device=cuda)
    o192 = ops.gdn_bwd_out(x192, x192.clone(), x192.clone(), sum_out=s192)
    assert _rel(s192, o192.float().reshape(-1, 192).double().sum(0).cpu()) <= 1e-4
    beta, gamma = torch.rand(16, generator=g) * 2e-3, torch.rand(16, 16, generator=g) * 8e-6
    db_hat, dg_hat = torch.randn(16, generator=g), torch.randn(16, 16, generator=g)
    bb, gb = 1.0004e-3, 3.8e-6
    got_b, got_g = ops.gdn_param_grad(beta.to(cuda), gamma.to(cuda), db_hat.to(cuda), dg_hat.to(cuda), bb, gb)
    for got, p_, d_, bound in ((got_b, beta, db_hat, bb), (got_g, gamma, dg_hat, gb)):
        dl = d_ * 2 * torch.clamp(p_, min=bound)
        ref = dl * ((p_ >= bound) | (dl < 0)).float()
        assert torch.allclose(got.cpu(), ref, rtol=1e-6, atol=0)
    ref = _bf(d_direct.float().cpu() + 2 * xf * t.float())
    assert _rel(dx.float(), ref.float()) <= 1e-2
    y = _bf(torch.randn(n, generator=g))
    r = ops.relu_bwd(y.to(cuda), gr.to(cuda))
    assert torch.equal(r.cpu(), torch.where(y > 0, gr, torch.zeros_like(gr)))
    cs = ops.colsum_bf16(x.to(cuda).reshape(-1, 128))
    assert _rel(cs, x.float().reshape(-1, 128).double().sum(0)) <= 1e-4
    x320 = x[:320 * 397].reshape(-1, 320)
    cs = ops.colsum_bf16(x320.to(cuda).contiguous())
    assert _rel(cs, x320.float().double().sum(0)) <= 1e-4


def test_conv1x1_engine_and_im2col(cuda):
    g = torch.Generator().manual_seed(6)
    C = 128
    x = _bf(torch.randn(2, 24, 40, C, generator=g))
    w = torch.randn(C, C, 1, 1, generator=g) * 0.1
    b = torch.randn(C, generator=g)
    packed = ops.pack_conv_weight(w.to(cuda), _lib.CONV_1X1, C, C, _lib.LAYOUT_NHWC_BF16)
    y = ops.conv_forward(x.to(cuda), kind=_lib.CONV_1X1, epilogue=_lib.EPI_NONE, in_layout=_lib.LAYOUT_NHWC_BF16,
                         out_layout=_lib.LAYOUT_NHWC_BF16, in_c=C, out_c=C, weight=packed, bias=b.to(cuda))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2).double(), _bf(w).double(), b.double()).permute(0, 2, 3, 1)
    assert _rel(y.float(), ref) <= 1e-2
    img = torch.rand(2, 3, 20, 28, generator=g)
    rows = ops.im2col5x5s2(img.to(cuda))
    ref = F.unfold(img, 5, padding=2, stride=2).transpose(1, 2).reshape(2, 10, 14, 75)
    assert rows.shape == (2, 10, 14, 80)
    assert torch.equal(rows[..., :75].cpu(), _bf(ref))
    assert float(rows[..., 75:].abs().max()) == 0.0


@pytest.mark.parametrize("bands,cs,batch,h,w", [(3, 128, 2, 64, 64), (1, 128, 1, 48, 40), (3, 192, 2, 34, 52), (3, 64, 1, 19, 28),
                                               (1, 256, 1, 16, 96), (3, 128, 40, 128, 128)])
def test_edge_wgrad_fused_equals_patch_matrix_path(cuda, bands, cs, batch, h, w):
    """licos_conv_wgrad_image (im2col tile built in shared memory) against the float64 product of the SAME bf16 operands
    (the patch matrix of licos_im2col5x5s2, which the test above pins to F.unfold), ragged tiles, 1 / 3 bands, 1 / 2
    accumulator blocks, more tiles than SMs."""
    g = torch.Generator().manual_seed(bands * 1000 + cs + h)
    oh, ow = (h + 1) // 2, (w + 1) // 2
    img = torch.rand(batch, bands, h, w, generator=g).to(cuda)
    small = torch.randn(batch, oh, ow, cs, generator=g).bfloat16().to(cuda)
    kp = ops.im2col_kpad(bands)
    out = torch.zeros(cs * kp, dtype=torch.float32, device=cuda)
    got = ops.conv_wgrad_image(small, img, out)
    rows = ops.im2col5x5s2(img)
    ref = (small.reshape(-1, cs).double().t() @ rows.reshape(-1, kp).double()).cpu()
    assert got.shape == (cs, kp)
    assert _rel(got, ref) <= 2e-5
    assert float(got[:, bands * 25:].abs().max()) == 0.0
    ops.conv_wgrad_image(small, img, out)  # accumulates
    assert _rel(out.view(cs, kp), 2 * ref) <= 2e-5


def test_edge_wgrad_fused_refuses_what_it_is_not_built_for(cuda):
    small = torch.zeros(1, 8, 8, 128, dtype=torch.bfloat16, device=cuda)
    with pytest.raises(NotImplementedError):  # 13 bands: the patch-matrix path
        ops.conv_wgrad_image(small, torch.zeros(1, 13, 16, 16, device=cuda), torch.zeros(128 * ops.im2col_kpad(13), device=cuda))
    with pytest.raises(NotImplementedError):  # rows TMA cannot load (width not a multiple of 4)
        ops.conv_wgrad_image(small, torch.zeros(1, 3, 16, 15, device=cuda), torch.zeros(128 * 80, device=cuda))


# ---------------------------------------------------------------------------------------------
# chain level
# ---------------------------------------------------------------------------------------------

def _pair(name, in_ch, quality=1):
    torch.manual_seed(42)
    if in_ch == 3:
        net = L.image_models[name](quality=quality, pretrained=False)
        ref = R.image_models[name](quality=quality)
    else:
        net = L.get_model(name, False, in_ch, quality)
        ref = R.get_model(name, False, in_ch, quality)
    synth.condition_weights(net)
    ref.load_state_dict(net.state_dict())
    return net, ref


def _grad_report(what, got, ref):
    got, ref = got.double().cpu().flatten(), ref.double().flatten()
    cos = float(torch.dot(got, ref) / (got.norm() * ref.norm() + 1e-300))
    rel = float((got - ref).norm() / (ref.norm() + 1e-300))
    print(f"grad {what}: cos {cos:.6f} rel-L2 {rel:.3e} |ref| {float(ref.norm()):.3e}")
    return cos, rel


def _check_chain(cuda, chain, ref_chain, x, what, x_grad=False, cos_min=0.999, rel_max=3e-2):
    g = torch.Generator().manual_seed(77)
    chain = chain.to(cuda).train()
    ref_chain = ref_chain.train()
    xr = x.clone().requires_grad_(x_grad)
    yr = ref_chain(xr)
    r = torch.randn(yr.shape, generator=g)
    (yr * r).sum().backward()
    xd = x.to(cuda).requires_grad_(x_grad)
    yd = chain(xd)
    assert yd.grad_fn is not None and "ChainFn" in type(yd.grad_fn).__name__, "native training path not taken"
    (yd * r.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    cos, rel = _grad_report(f"{what} output", yd.detach(), yr.detach())
    assert cos >= cos_min
    worst = (1.0, 0.0)
    for (n, p), (_, q) in zip(chain.named_parameters(), ref_chain.named_parameters()):
        assert p.grad is not None, n
        assert p.grad.shape == q.grad.shape
        c, e = _grad_report(f"{what} {n}", p.grad, q.grad)
        worst = (min(worst[0], c), max(worst[1], e))
        assert c >= cos_min and e <= rel_max, n
    if x_grad:
        c, e = _grad_report(f"{what} input", xd.grad, xr.grad)
        assert c >= cos_min and e <= rel_max
    print(f"{what}: worst cos {worst[0]:.6f} worst rel-L2 {worst[1]:.3e}")


@pytest.mark.parametrize("in_ch", [3, 1])
def test_g_a_gradients_vs_oracle(cuda, in_ch):
    net, ref = _pair("bmshj2018-factorized", in_ch)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, in_ch, 128, 128, generator=g)
    _check_chain(cuda, net.g_a, ref.g_a, x, f"g_a c{in_ch}")


def test_g_s_gradients_vs_oracle(cuda):
    net, ref = _pair("bmshj2018-factorized", 3)
    g = torch.Generator().manual_seed(2)
    y = torch.round(torch.randn(2, 192, 8, 8, generator=g) * 3)
    _check_chain(cuda, net.g_s, ref.g_s, y, "g_s", x_grad=True)


def test_hyperprior_chains_gradients_vs_oracle(cuda):
    net, ref = _pair("bmshj2018-hyperprior", 3)
    g = torch.Generator().manual_seed(3)
    y = torch.randn(2, 192, 16, 16, generator=g).abs() * 2
    _check_chain(cuda, net.h_a, ref.h_a, y, "h_a", x_grad=True, cos_min=0.995, rel_max=0.1)
    z = torch.round(torch.randn(2, 128, 4, 4, generator=g) * 2)
    _check_chain(cuda, net.h_s, ref.h_s, z, "h_s", x_grad=True, cos_min=0.995, rel_max=0.1)


def test_relu_variant_gradients_vs_oracle(cuda):
    net, ref = _pair("bmshj2018-factorized-relu", 3)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 64, 64, generator=g)
    _check_chain(cuda, net.g_a, ref.g_a, x, "g_a relu", cos_min=0.99, rel_max=0.15)


def test_train_step_loss_and_grads_vs_oracle(cuda):
    """One train.py-style step (forward in train mode with a fixed noise tensor is not reachable through the model API,
    so the rate term is compared in eval mode: rounding instead of noise): loss and every gradient."""
    net, ref = _pair("bmshj2018-factorized", 3)
    net = net.to(cuda)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(2, 3, 128, 128, generator=g)
    crit, rcrit = L.RateDistortionLoss(lmbda=1e-2), R.RateDistortionLoss(lmbda=1e-2)
    net.g_a.train(); net.g_s.train(); ref.train()
    ref.entropy_bottleneck.eval(); net.entropy_bottleneck.eval()
    out = net(x.to(cuda))
    loss = crit(out, x.to(cuda))["loss"]
    loss.backward()
    rout = ref(x)
    rloss = rcrit(rout, x)["loss"]
    rloss.backward()
    print(f"loss {float(loss):.5f} vs oracle {float(rloss):.5f}")
    assert abs(float(loss) / float(rloss) - 1) <= 5e-3
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        if q.grad is None or float(q.grad.norm()) == 0:
            continue
        assert p.grad is not None, n
        c, e = _grad_report(n, p.grad, q.grad)
        if n.startswith("g_s"):  # g_a's gradients pass through the rounding of y: different symbols, different gradient
            assert c >= 0.995 and e <= 0.1, n


@pytest.mark.parametrize("in_ch,form", [(3, "plain"), (3, "stable"), (1, "plain"), (13, "plain")])
def test_entropy_bottleneck_training_backward_vs_oracle(cuda, in_ch, form):
    """EntropyBottleneck.forward(training=True) with a given noise tensor: y_hat / likelihoods and the gradients of a
    bpp-like loss + a linear term on y_hat with respect to x and every density parameter (fp32 on both sides)."""
    net, ref = _pair("bmshj2018-factorized", in_ch)
    eb, reb = net.entropy_bottleneck.to(cuda).train(), ref.entropy_bottleneck.train()
    eb.likelihood_form = reb.likelihood_form = form
    g = torch.Generator().manual_seed(21)
    C = eb.channels
    x = torch.randn(3, C, 6, 10, generator=g) * 4
    x[0, 0, 0, :4] = torch.tensor([60.0, -70.0, 45.0, -48.0])  # likelihood below the 1e-9 bound: LowerBound's gradient rule
    noise = torch.rand(x.shape, generator=g) - 0.5
    r = torch.randn(x.shape, generator=g)

    def run(m, dev):
        xi = x.to(dev).requires_grad_(True)
        y_hat, lik = m(xi, training=True, noise=noise.to(dev))
        loss = -torch.log(lik).sum() / 7.0 + (y_hat * r.to(dev)).sum()
        loss.backward()
        return y_hat.detach().cpu(), lik.detach().cpu(), xi.grad.cpu(), loss.item()

    yh, lk, gx, loss = run(eb, cuda)
    ryh, rlk, rgx, rloss = run(reb, "cpu")
    assert type(eb(x.to(cuda).requires_grad_(True), training=True)[0].grad_fn).__name__.startswith("_EbTrainFn")
    assert torch.allclose(yh, ryh, atol=1e-6)
    assert _rel(lk, rlk) <= 1e-4
    c, e = _grad_report(f"eb c{in_ch} {form} input", gx, rgx)
    assert c >= 0.99999 and e <= 2e-3
    for (n, p), (_, q) in zip(eb.named_parameters(), reb.named_parameters()):
        if q.grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0, n
            continue
        c, e = _grad_report(f"eb c{in_ch} {form} {n}", p.grad, q.grad)
        assert c >= 0.9999 and e <= 5e-3, n


def test_graphed_train_step_matches_eager(cuda):
    """GraphedTrainStep replays the same step the eager loop runs: same loss trajectory with the noise fixed by the
    generator seed, parameters move, and the inference path sees the updated weights afterwards."""
    def make():
        torch.manual_seed(5)
        net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
        synth.condition_weights(net)
        net = net.to(cuda).train()
        opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
        return net, opt

    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 3, 64, 64, generator=g).to(cuda)
    crit = L.RateDistortionLoss(lmbda=1e-2)
    net, opt = make()
    w0 = net.g_a[0].weight.detach().clone()
    step = L.GraphedTrainStep(net, crit, opt, x, clip_max_norm=1.0, warmup=2)
    losses = [float(step(x)["loss"]) for _ in range(6)]
    assert all(l == l for l in losses)
    assert losses[-1] < losses[0], losses                      # it trains
    assert float((net.g_a[0].weight - w0).abs().max()) > 0
    net.eval()
    with torch.no_grad():
        a = net(x)["x_hat"]
        net.g_a._packed_cache.clear(); net.g_s._packed_cache.clear()   # force a repack from the current parameters
        b = net(x)["x_hat"]
    assert torch.equal(a, b), "inference after graphed training used stale packed weights"
    # eager loop, same number of steps (2 warm-up + 6 replays = 8 optimizer steps; capture itself executes nothing)
    net2, opt2 = make()
    for _ in range(8):
        opt2["net"].zero_grad(); opt2["aux"].zero_grad()
        out = net2(x)
        l2 = crit(out, x)["loss"]
        l2.backward()
        torch.nn.utils.clip_grad_norm_(net2.parameters(), 1.0)
        opt2["net"].step()
        aux = net2.aux_loss(); aux.backward(); opt2["aux"].step()
    print(f"loss at step 8: graphed {losses[-1]:.4f} eager {float(l2):.4f}")
    assert abs(losses[-1] / float(l2) - 1) < 0.05  # different noise draws and red.add order, same trajectory


@pytest.mark.parametrize("inverse,pixels", [(False, 3 * 37 * 29), (True, 3 * 37 * 29), (False, 50000), (True, 128 * 148 * 4)])
def test_gdn_backward_fused_kernel(cuda, inverse, pixels):
    """licos_gdn_backward (one pass, 128 channels) against the formulas in float64 on the same bf16 inputs; the kernel
    rounds x^2, d_norm and dx to bf16 where the reference keeps float64, hence the 1e-2 scale-relative bar."""
    g = torch.Generator().manual_seed(pixels % 1000 + int(inverse))
    C = 128
    x = _bf(torch.randn(pixels, C, generator=g) * 1.5)
    gr = _bf(torch.randn(pixels, C, generator=g))
    gamma = _bf(torch.rand(C, C, generator=g) * 0.02 + 0.1 * torch.eye(C))
    beta = torch.rand(C, generator=g) + 0.5
    d_gamma = torch.zeros(C, C, device=cuda)
    d_beta = torch.zeros(C, device=cuda)
    d_bias = torch.zeros(C, device=cuda)
    dx = ops.gdn_backward(x.to(cuda), gr.to(cuda), gamma.to(cuda), beta.to(cuda), inverse, d_gamma, d_beta, d_bias)
    torch.cuda.synchronize()
    xd, gd, gm, bt = x.double(), gr.double(), gamma.double(), beta.double()
    x2 = _bf((xd * xd).float()).double()
    n = bt + x2 @ gm.t()
    if inverse:
        dd, dn = gd * n.sqrt(), 0.5 * gd * xd / n.sqrt()
    else:
        dd, dn = gd * n.rsqrt(), -0.5 * gd * xd * n.pow(-1.5)
    dn_b = _bf(dn.float()).double()
    ref_dx = dd + 2 * xd * (dn_b @ gm)
    errs = {"dx": _rel(dx.float(), ref_dx), "d_gamma": _rel(d_gamma, dn_b.t() @ x2), "d_beta": _rel(d_beta, dn_b.sum(0)),
            "d_bias": _rel(d_bias, _bf(ref_dx.float()).double().sum(0))}
    print(f"gdn_backward fused inverse={inverse} P={pixels}: max|err|/max = " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert errs["dx"] <= 1e-2 and errs["d_gamma"] <= 1e-2 and errs["d_beta"] <= 1e-2 and errs["d_bias"] <= 2e-2


@pytest.mark.parametrize("in_ch", [3, 13])
def test_entropy_bottleneck_native_pack_and_aux_loss(cuda, in_ch):
    """licos_eb_pack_params against the torch packing, and EntropyBottleneck.loss() (aux loss) + its quantile gradient
    against the oracle."""
    from licos_b200.entropy_models import _native_packed
    net, ref = _pair("bmshj2018-factorized", in_ch)
    eb, reb = net.entropy_bottleneck.to(cuda).train(), ref.entropy_bottleneck.train()
    a = eb.packed_params(force=True).packed
    b = _native_packed(eb, eb._params()[:-1]).packed
    assert _rel(b, a.cpu()) <= 1e-6
    la, lr = eb.loss(), reb.loss()
    assert "EbAuxLossFn" in type(la.grad_fn).__name__
    la.backward(); lr.backward()
    print(f"aux loss {float(la):.4f} vs oracle {float(lr):.4f}")
    assert abs(float(la) / float(lr) - 1) <= 1e-5
    assert _rel(eb.quantiles.grad, reb.quantiles.grad) <= 1e-4
    for n, p in eb.named_parameters():
        if not n.endswith("quantiles"):
            assert p.grad is None, n


@pytest.mark.parametrize("name", ["bmshj2018-hyperprior", "bmshj2018-factorized-relu"])
def test_other_models_train_step_runs_native(cuda, name):
    """train.py's step on the other two architectures get_model accepts (model_utils.py:20-24): the four transforms take
    the native path, every parameter receives a finite gradient, and a few Adam steps lower the loss."""
    torch.manual_seed(3)
    net = L.image_models[name](quality=1, pretrained=False)
    synth.condition_weights(net)
    net = net.to(cuda).train()
    opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
    crit = L.RateDistortionLoss(lmbda=1e-2)
    x = torch.rand(4, 3, 128, 128, generator=torch.Generator().manual_seed(12)).to(cuda)
    losses = []
    for it in range(6):
        opt["net"].zero_grad(); opt["aux"].zero_grad()
        out = net(x)
        assert "ChainFn" in type(out["x_hat"].grad_fn).__name__
        loss = crit(out, x)["loss"]
        loss.backward()
        if it == 0:
            for n, p in net.named_parameters():
                if n.endswith("quantiles"):
                    continue
                assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt["net"].step()
        aux = net.aux_loss(); aux.backward(); opt["aux"].step()
        losses.append(float(loss))
    print(f"{name}: loss {losses[0]:.3f} -> {losses[-1]:.3f}")
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("with_means", [False, True])
def test_gaussian_conditional_training_backward_vs_oracle(cuda, with_means):
    """GaussianConditional.forward(training=True) with a given noise tensor: outputs and the gradients with respect to y,
    the scales (both sides of the 0.11 bound) and the means, against the oracle's autograd (fp32 on both sides)."""
    gc, rgc = L.GaussianConditional(None).to(cuda).train(), R.GaussianConditional(None).train()
    g = torch.Generator().manual_seed(31 + int(with_means))
    shape = (2, 8, 12, 10)
    y = torch.randn(shape, generator=g) * 3
    scales = torch.exp(torch.randn(shape, generator=g) * 1.5 - 1.0)      # 0.02 .. 20: both sides of the scale bound
    y[0, 0, 0, :3] = torch.tensor([90.0, -80.0, 70.0]); scales[0, 0, 0, :3] = 0.5   # likelihood below 1e-9
    means = torch.randn(shape, generator=g) if with_means else None
    noise = torch.rand(shape, generator=g) - 0.5
    r = torch.randn(shape, generator=g)

    def run(m, dev):
        yi, si = y.to(dev).requires_grad_(True), scales.to(dev).requires_grad_(True)
        mi = means.to(dev).requires_grad_(True) if with_means else None
        y_hat, lik = m(yi, si, mi, training=True, noise=noise.to(dev))
        (-torch.log(lik).sum() / 5.0 + (y_hat * r.to(dev)).sum()).backward()
        return (y_hat.detach().cpu(), lik.detach().cpu(), yi.grad.cpu(), si.grad.cpu(), mi.grad.cpu() if with_means else None,
                type(y_hat.grad_fn).__name__)

    yh, lk, gy, gs, gm, fn = run(gc, cuda)
    ryh, rlk, rgy, rgs, rgm, _ = run(rgc, "cpu")
    assert fn.startswith("_GcTrainFn")
    assert torch.allclose(yh, ryh, atol=1e-6) and _rel(lk, rlk) <= 1e-5
    for what, a, b in (("y", gy, rgy), ("scales", gs, rgs)) + ((("means", gm, rgm),) if with_means else ()):
        c, e = _grad_report(f"gc {what}", a, b)
        assert c >= 0.99999 and e <= 2e-3, what


def test_wide_model_gradients_vs_oracle(cuda):
    """Quality 6 (N = 192, M = 320): the 192-channel GDN backward takes the unfused sequence (square, 1x1 layers, mid / out
    kernels, 1x1 weight gradient), the weight-gradient kernel runs three channel blocks (320 = 128 + 128 + 64)."""
    torch.manual_seed(42)
    net = L.image_models["bmshj2018-factorized"](quality=6, pretrained=False)
    ref = R.image_models["bmshj2018-factorized"](quality=6)
    synth.condition_weights(net)
    ref.load_state_dict(net.state_dict())
    g = torch.Generator().manual_seed(13)
    x = torch.rand(2, 3, 64, 96, generator=g)
    _check_chain(cuda, net.g_a, ref.g_a, x, "g_a q6")
    y = torch.round(torch.randn(2, 320, 4, 6, generator=g) * 3)
    _check_chain(cuda, net.g_s, ref.g_s, y, "g_s q6", x_grad=True)


@pytest.mark.parametrize("in_ch", [3, 13])
def test_g_a_input_gradient_vs_oracle(cuda, in_ch):
    """d loss / d image through g_a (the data gradient of the first layer: a 128 -> C_img transposed conv, fp32 NCHW out)."""
    net, ref = _pair("bmshj2018-factorized", in_ch)
    g = torch.Generator().manual_seed(17)
    x = torch.rand(2, in_ch, 64, 64, generator=g)
    _check_chain(cuda, net.g_a, ref.g_a, x, f"g_a c{in_ch} with input gradient", x_grad=True)


def test_wide_bottlenecks_and_odd_sizes_are_refused_at_forward_time(cuda):
    """No detour through torch autograd: shapes the backward kernels do not take raise NotImplementedError when the forward
    is called (not from inside autograd's backward in the middle of a step)."""
    net = L.get_model("bmshj2018-factorized", False, 16, 1).to(cuda).train()   # filters (16, 16, 3, 3): 409 parameters / channel
    x = torch.rand(1, 16, 64, 64, device=cuda)
    with pytest.raises(NotImplementedError, match="density parameters"):
        net(x)
    with torch.no_grad():                                                        # inference on the same model is fine
        assert net.eval()(x)["x_hat"].shape == x.shape
    net3 = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False).to(cuda).train()
    with pytest.raises(NotImplementedError, match="even sizes"):
        net3(torch.rand(1, 3, 66, 70, device=cuda))                              # 66 -> 33: odd under a stride-2 layer
