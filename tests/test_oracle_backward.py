"""CPU checks (float64, torch autograd on the oracle's modules) of the identities the native backward kernels are built on
-- the counterpart of test_oracle_model.py's forward decompositions.  If one of these fails the kernels compute the wrong
thing no matter how well they match their own formulas.

  * data gradient of Conv2d(5, s2, p2)          == ConvTranspose2d(5, s2, p2, op1) over the output gradient, SAME weight
  * data gradient of ConvTranspose2d(5,s2,p2,1) == Conv2d(5, s2, p2) over the output gradient, SAME weight
  * data gradient of Conv2d(3, s1, p1)          == Conv2d(3, s1, p1) with the flipped, transposed weight
  * weight gradients == sum over pixels of (small tensor) x (big tensor shifted by the tap)        [licos_conv_wgrad]
  * GDN / IGDN backward == the norm / d_norm / t / dx sequence of licos_gdn_backward, incl. gamma, beta gradients
  * GaussianConditional backward == phi at the two bin edges                                       [licos_gc_backward]
  * EntropyBottleneck backward == per-element reverse sweep through the packed parameters   [licos_eb_backward + param_grads]
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import compressai_ref as R


def _close(a, b, tol=1e-10):
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def test_conv_data_gradient_is_the_transposed_conv_with_the_same_weight():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 12, 16, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(7, 5, 5, 5, generator=g, dtype=torch.float64)
    y = F.conv2d(x, w, stride=2, padding=2)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * gy).sum().backward()
    # torch reads a ConvTranspose2d weight as (in, out, kh, kw): w's (O, I, 5, 5) is exactly that with in = O
    assert _close(F.conv_transpose2d(gy, w, stride=2, padding=2, output_padding=1), x.grad)


def test_deconv_data_gradient_is_the_conv_with_the_same_weight():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 6, 7, 5, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(6, 4, 5, 5, generator=g, dtype=torch.float64)  # ConvTranspose2d weight (in, out, kh, kw)
    y = F.conv_transpose2d(x, w, stride=2, padding=2, output_padding=1)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * gy).sum().backward()
    assert _close(F.conv2d(gy, w, stride=2, padding=2), x.grad)   # read as a Conv2d weight (out = 6, in = 4)


def test_conv3x3_data_gradient_is_the_flipped_transposed_conv():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 9, 6, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(5, 4, 3, 3, generator=g, dtype=torch.float64)
    y = F.conv2d(x, w, stride=1, padding=1)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * gy).sum().backward()
    assert _close(F.conv2d(gy, w.flip(2, 3).transpose(0, 1), stride=1, padding=1), x.grad)


def _wgrad_formula(small, big, k, stride, pad):
    """out[kh*k + kw][cs][cb] = sum_{b,i,j} small[b,cs,i,j] * big[b,cb, s*i + kh - pad, s*j + kw - pad] (zero outside)."""
    B, cs, h, w = small.shape
    cb = big.shape[1]
    bp = F.pad(big, (pad, pad + stride, pad, pad + stride))
    out = torch.zeros(k * k, cs, cb, dtype=small.dtype)
    for kh in range(k):
        for kw in range(k):
            view = bp[:, :, kh:kh + stride * h:stride, kw:kw + stride * w:stride]
            out[kh * k + kw] = torch.einsum("bshw,bchw->sc", small, view)
    return out


@pytest.mark.parametrize("transposed", [False, True])
def test_weight_gradient_is_a_contraction_over_pixels(transposed):
    g = torch.Generator().manual_seed(4 + int(transposed))
    if not transposed:   # Conv2d: small = output gradient, big = input
        x = torch.randn(2, 3, 10, 8, generator=g, dtype=torch.float64)
        w = torch.randn(4, 3, 5, 5, generator=g, dtype=torch.float64, requires_grad=True)
        y = F.conv2d(x, w, stride=2, padding=2)
        gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        (y * gy).sum().backward()
        out = _wgrad_formula(gy, x, 5, 2, 2)                       # [25][O][I]
        assert _close(out.permute(1, 2, 0).reshape(4, 3, 5, 5), w.grad)
    else:                # ConvTranspose2d: small = input, big = output gradient
        x = torch.randn(2, 3, 5, 4, generator=g, dtype=torch.float64)
        w = torch.randn(3, 4, 5, 5, generator=g, dtype=torch.float64, requires_grad=True)
        y = F.conv_transpose2d(x, w, stride=2, padding=2, output_padding=1)
        gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        (y * gy).sum().backward()
        out = _wgrad_formula(x, gy, 5, 2, 2)                       # [25][I][O]
        assert _close(out.permute(1, 2, 0).reshape(3, 4, 5, 5), w.grad)


def test_conv3x3_weight_gradient_formula():
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 3, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, stride=1, padding=1)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (y * gy).sum().backward()
    assert _close(_wgrad_formula(gy, x, 3, 1, 1).permute(1, 2, 0).reshape(4, 3, 3, 3), w.grad)


@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_backward_sequence(inverse):
    """The sequence of licos_gdn_backward against autograd through the oracle's GDN (reparametrised gamma / beta and
    LowerBound's rule included: licos_gdn_param_grad)."""
    torch.manual_seed(7)
    C = 6
    gdn = R.GDN(C, inverse=inverse).double()
    with torch.no_grad():
        gdn.gamma.add_(torch.rand(C, C, dtype=torch.float64) * 0.05)
        gdn.gamma[0, 1] = 1e-7      # below the gamma bound (2^-18): LowerBound's rule decides whether a gradient passes
        gdn.beta[2] = 5e-4          # below the beta bound
    x = torch.randn(2, C, 5, 4, dtype=torch.float64, requires_grad=True)
    y = gdn(x)
    g = torch.randn(y.shape, dtype=torch.float64)
    (y * g).sum().backward()

    xv, gv = x.detach().permute(0, 2, 3, 1).reshape(-1, C), g.permute(0, 2, 3, 1).reshape(-1, C)
    lb_b = torch.max(gdn.beta.detach(), gdn.beta_reparam.lower_bound.bound)
    lb_g = torch.max(gdn.gamma.detach(), gdn.gamma_reparam.lower_bound.bound)
    beta_hat, gamma_hat = lb_b ** 2 - gdn.beta_reparam.pedestal, lb_g ** 2 - gdn.gamma_reparam.pedestal
    norm = beta_hat + (xv * xv) @ gamma_hat.t()                                  # GEMM 1
    if inverse:
        d_direct, d_norm = gv * norm.sqrt(), 0.5 * gv * xv / norm.sqrt()
    else:
        d_direct, d_norm = gv * norm.rsqrt(), -0.5 * gv * xv * norm.pow(-1.5)
    t = d_norm @ gamma_hat                                                        # GEMM 2
    dx = d_direct + 2 * xv * t
    d_gamma_hat = d_norm.t() @ (xv * xv)                                          # GEMM 3
    d_beta_hat = d_norm.sum(0)
    assert _close(dx.reshape(2, 5, 4, C).permute(0, 3, 1, 2), x.grad)

    def param_grad(p, bound, lb, d_hat):                                          # licos_gdn_param_grad
        d_lb = d_hat * 2 * lb
        return d_lb * ((p >= bound) | (d_lb < 0)).double()
    assert _close(param_grad(gdn.gamma.detach(), gdn.gamma_reparam.lower_bound.bound, lb_g, d_gamma_hat), gdn.gamma.grad)
    assert _close(param_grad(gdn.beta.detach(), gdn.beta_reparam.lower_bound.bound, lb_b, d_beta_hat), gdn.beta.grad)


@pytest.mark.parametrize("with_means", [False, True])
def test_gaussian_conditional_backward_formula(with_means):
    g = torch.Generator().manual_seed(8)
    gc = R.GaussianConditional(None).double()
    shape = (2, 3, 4, 5)
    y = (torch.randn(shape, generator=g, dtype=torch.float64) * 3).requires_grad_(True)
    scales = torch.exp(torch.randn(shape, generator=g, dtype=torch.float64) * 1.5 - 1.0).requires_grad_(True)
    means = torch.randn(shape, generator=g, dtype=torch.float64).requires_grad_(True) if with_means else None
    noise = torch.rand(shape, generator=g, dtype=torch.float64) - 0.5
    y_hat, lik = gc(y, scales, means, training=True, noise=noise)
    g_lik = torch.randn(shape, generator=g, dtype=torch.float64)
    (lik * g_lik).sum().backward()

    bound = float(gc.lower_bound_scale.bound)
    diff = (y_hat.detach() - (means.detach() if with_means else 0.0))
    v, s = diff.abs(), torch.clamp(scales.detach(), min=bound)
    a, b = (0.5 - v) / s, (-0.5 - v) / s
    phi = lambda z: torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)  # noqa: E731
    lik_raw = 0.5 * torch.erfc(-a / math.sqrt(2)) - 0.5 * torch.erfc(-b / math.sqrt(2))
    gg = torch.where((lik_raw >= 1e-9) | (g_lik < 0), g_lik, torch.zeros_like(g_lik))
    d_y = gg * (phi(b) - phi(a)) / s * torch.sign(diff)
    d_s = gg * (b * phi(b) - a * phi(a)) / s
    d_s = torch.where((scales.detach() >= bound) | (d_s < 0), d_s, torch.zeros_like(d_s))
    assert _close(d_y, y.grad) and _close(d_s, scales.grad)
    if with_means:
        assert _close(-d_y, means.grad)


@pytest.mark.parametrize("form", ["plain", "stable"])
def test_entropy_bottleneck_backward_sweep(form):
    """licos_eb_backward + licos_eb_param_grads restated per element in float64: forward through the packed parameters
    (softplus / tanh applied), reverse sweep with the activations kept, LowerBound gating on the likelihood, then the chain
    back to the raw parameters -- against autograd through the oracle's EntropyBottleneck."""
    torch.manual_seed(9)
    C = 3
    eb = R.EntropyBottleneck(C).double()
    eb.likelihood_form = form
    with torch.no_grad():
        for n, p in eb.named_parameters():
            if "factor" in n:
                p.normal_(0, 0.5)
            elif "matrix" in n:
                p.add_(torch.randn_like(p) * 0.3)
    x = (torch.randn(2, C, 3, 4, dtype=torch.float64) * 4)
    x[0, 0, 0, :2] = torch.tensor([60.0, -70.0])     # likelihood under the 1e-9 bound
    x.requires_grad_(True)
    noise = torch.rand(x.shape, dtype=torch.float64) - 0.5
    y_hat, lik = eb(x, training=True, noise=noise)
    g_lik = torch.randn(x.shape, dtype=torch.float64)
    g_lik[0, 0, 0, 0] = -1.0                          # negative gradient passes the bound, positive does not
    g_lik[0, 0, 0, 1] = 1.0
    g_yhat = torch.randn(x.shape, dtype=torch.float64)
    ((lik * g_lik).sum() + (y_hat * g_yhat).sum()).backward()

    n_layers = len(eb.filters) + 1
    mats = [F.softplus(getattr(eb, f"_matrix{i}").detach()) for i in range(n_layers)]
    bias = [getattr(eb, f"_bias{i}").detach() for i in range(n_layers)]
    facs = [torch.tanh(getattr(eb, f"_factor{i}").detach()) for i in range(n_layers - 1)]
    d_m = [torch.zeros_like(m) for m in mats]
    d_b = [torch.zeros_like(b) for b in bias]
    d_f = [torch.zeros_like(f) for f in facs]
    d_x = torch.zeros_like(x)
    sig = torch.sigmoid
    for b_ in range(x.shape[0]):
        for c in range(C):
            for i_ in range(x.shape[2]):
                for j_ in range(x.shape[3]):
                    v = y_hat[b_, c, i_, j_].detach()

                    def fwd(inp):
                        cur, ins, ths = inp.reshape(1, 1), [], []
                        for L in range(n_layers):
                            ins.append(cur)
                            a = mats[L][c] @ cur + bias[L][c]
                            th = torch.tanh(a) if L < n_layers - 1 else torch.zeros_like(a)
                            ths.append(th)
                            cur = a + facs[L][c] * th if L < n_layers - 1 else a
                        return cur.reshape(()), ins, ths

                    def bwd(seed, ins, ths):
                        d_out = seed.reshape(1, 1)
                        for L in range(n_layers - 1, -1, -1):
                            if L < n_layers - 1:
                                d_f[L][c] += d_out * ths[L]
                                da = d_out * (1 + facs[L][c] * (1 - ths[L] ** 2))
                            else:
                                da = d_out
                            d_b[L][c] += da
                            d_m[L][c] += da @ ins[L].t()
                            d_out = mats[L][c].t() @ da
                        return d_out.reshape(())

                    lower, ins_l, ths_l = fwd(v - 0.5)
                    upper, ins_u, ths_u = fwd(v + 0.5)
                    if form == "plain":
                        a, bq = sig(upper), sig(lower)
                        likv, su, sl = a - bq, a * (1 - a), -bq * (1 - bq)
                    else:
                        s = -torch.sign(lower + upper)
                        a, bq = sig(s * upper), sig(s * lower)
                        likv = (a - bq).abs()
                        sg = torch.sign(a - bq)
                        su, sl = sg * s * a * (1 - a), -sg * s * bq * (1 - bq)
                    gl = g_lik[b_, c, i_, j_]
                    if not (likv >= 1e-9 or gl < 0):
                        gl = gl * 0
                    d_x[b_, c, i_, j_] = bwd(gl * su, ins_u, ths_u) + bwd(gl * sl, ins_l, ths_l) + g_yhat[b_, c, i_, j_]
    assert _close(d_x, x.grad, 1e-9)
    for L in range(n_layers):
        raw_m = getattr(eb, f"_matrix{L}")
        assert _close(d_m[L] * torch.sigmoid(raw_m.detach()), raw_m.grad, 1e-9)          # d softplus = sigmoid
        assert _close(d_b[L], getattr(eb, f"_bias{L}").grad, 1e-9)
        if L < n_layers - 1:
            raw_f = getattr(eb, f"_factor{L}")
            assert _close(d_f[L] * (1 - torch.tanh(raw_f.detach()) ** 2), raw_f.grad, 1e-9)
