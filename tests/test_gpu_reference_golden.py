"""-m gpu: the CUDA path against tests/golden/reference_glue.npz -- numbers produced by executing the REFERENCE'S OWN
glue files (eval_utils.py process_img / compute_bpp / compute_psnr / compute_msssim, licos/train.py train_one_batch,
licos/model_utils.py get_model, licos/raw_image_folder.py _open_band_) over the CPU oracle in the build container
(tests/golden/make_reference_golden.py; tests/test_reference_glue.py keeps the fixture equal to what that code produces).
The reference tree is not on the GPU box, so the few lines of glue are restated here next to their file:line.

Tolerances (BASELINE.json north_star, bf16 tensor-core path): PSNR within 0.01 dB, bpp within 0.1 %."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reference_cases as RC  # noqa: E402

import licos_b200 as L  # noqa: E402
from licos_b200 import ops  # noqa: E402
from oracle import compressai_ref as R  # noqa: E402  (checker: rebuilds the seeded weights the fixture was made with)

pytestmark = pytest.mark.gpu
GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_glue.npz"))
PSNR_TOL_DB, BPP_TOL_REL = 0.01, 1e-3


def _product_with_fixture_weights(case, fingerprint, device):
    model, c, q = case
    ref = RC.build_weights(R.get_model, case)
    assert np.allclose(RC.state_fingerprint(ref.state_dict()), fingerprint, rtol=1e-12, atol=0), "weight recipe drifted"
    net = L.get_model(model, False, c, q)
    net.load_state_dict(ref.state_dict())
    return net.to(device)


@pytest.mark.parametrize("name", list(RC.EVAL_CASES))
def test_process_img_against_reference_run(cuda, name):
    g = {k.split("/", 2)[2]: GOLDEN[k] for k in GOLDEN.files if k.startswith(f"eval/{name}/")}
    net = _product_with_fixture_weights(RC.EVAL_CASES[name][:3], g["fingerprint"], cuda).eval()
    net.update()
    img = RC.eval_image(name).to(cuda)
    # eval_utils.py:199-210 (process_img)
    with torch.no_grad():
        out_net = net.forward(img.unsqueeze(0))
        compressed = net.compress(img.unsqueeze(0))
    nbytes = np.frombuffer(np.array(compressed["strings"]), dtype=np.uint8).size
    out_net["x_hat"].clamp_(0, 1)
    out_net["x_hat"] = out_net["x_hat"][..., : img.shape[1], : img.shape[2]]
    diff = torch.mean((out_net["x_hat"] - img).abs(), axis=1).squeeze().cpu()
    x_hat = out_net["x_hat"]
    assert tuple(x_hat.shape) == tuple(g["x_hat_shape"])
    assert [list(v.shape) for v in out_net["likelihoods"].values()] == g["lik_shapes"].tolist()
    bpp = L.compute_bpp(out_net)                                           # eval_utils.py:172-186
    psnr = L.compute_psnr(img.unsqueeze(0), x_hat.contiguous())            # eval_utils.py:145-156
    print(f"{name}: PSNR {psnr:.5f} vs {float(g['psnr']):.5f} dB, bpp {bpp:.6f} vs {float(g['bpp']):.6f} "
          f"({(bpp / float(g['bpp']) - 1) * 100:+.4f} %), bytes {nbytes} vs {int(g['bytes'])}")
    assert abs(psnr - float(g["psnr"])) <= PSNR_TOL_DB
    assert abs(bpp / float(g["bpp"]) - 1) <= BPP_TOL_REL
    assert abs(nbytes / int(g["bytes"]) - 1) <= 0.01
    assert abs(float(diff.double().mean()) - float(g["diff_mean"])) <= 2e-3
    low = torch.nn.functional.adaptive_avg_pool2d(x_hat.float().cpu(), (8, 8)).numpy()
    assert np.abs(low - g["x_hat_lowres"]).max() <= 5e-3
    if "msssim" in g:
        ms = L.compute_msssim(img.unsqueeze(0), x_hat.contiguous())        # eval_utils.py:159-169
        assert abs(ms - float(g["msssim"])) <= 2e-3, (ms, float(g["msssim"]))
    # the strings really decode to the image forward() returned (decompress is UPSTREAM API, unused by LICOS)
    with torch.no_grad():
        dec = net.decompress(compressed["strings"], compressed["shape"])["x_hat"]
    assert torch.equal(dec[..., : img.shape[1], : img.shape[2]], x_hat)


def test_train_one_batch_against_reference_run(cuda):
    """licos/train.py:186-200 with the noise the reference run drew: losses, every gradient norm, the stored gradients."""
    g = {k.split("/", 1)[1]: GOLDEN[k] for k in GOLDEN.files if k.startswith("train/")}
    model, c, q, shape, _, _ = RC.TRAIN_CASE
    ref = RC.build_weights(R.get_model, (model, c, q))
    net = L.get_model(model, False, c, q)
    net.load_state_dict(ref.state_dict())
    net = net.to(cuda)
    conf = {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}}   # utils.py:65-73
    opt = L.net_aux_optimizer(net, conf)
    optimizer, aux_optimizer = opt["net"], opt["aux"]
    assert sum(len(gr["params"]) for gr in optimizer.param_groups) == int(g["n_net_params"])
    assert sum(len(gr["params"]) for gr in aux_optimizer.param_groups) == int(g["n_aux_params"])
    criterion = L.RateDistortionLoss(lmbda=1e-2)
    d = RC.train_inputs().to(cuda)
    noise = RC.train_noise((shape[0], net.entropy_bottleneck.channels, shape[2] // 16, shape[3] // 16)).to(cuda)
    eb_forward = net.entropy_bottleneck.forward
    net.entropy_bottleneck.forward = lambda x, **kw: eb_forward(x, noise=noise, **kw)  # pin the random draw only
    before = {k: v.detach().clone() for k, v in net.named_parameters()}

    net.train()                                                    # train.py:175
    optimizer.zero_grad()
    aux_optimizer.zero_grad()
    out_net = net(d)                                               # :190
    out_criterion = criterion(out_net, d)                          # :192
    out_criterion["loss"].backward()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)          # :194-195
    optimizer.step()
    aux_loss = net.aux_loss()                                      # :198
    aux_loss.backward()
    aux_optimizer.step()

    for k in ("loss", "mse_loss", "bpp_loss"):
        got, want = float(out_criterion[k]), float(g[k])
        print(f"{k}: {got:.6f} vs reference run {want:.6f}")
        assert abs(got / want - 1) <= 3e-3, k
    assert abs(float(net.aux_loss()) / float(g["aux_loss"]) - 1) <= 1e-3
    grads = {k: p.grad for k, p in net.named_parameters()}
    keys = [str(k) for k in g["grad_keys"]]
    assert sorted(grads) == keys
    worst = 0.0
    for k, want in zip(keys, g["grad_norms"]):
        got = float(grads[k].double().norm())
        worst = max(worst, abs(got / want - 1))
        assert abs(got / want - 1) <= 0.03, (k, got, want)
    print(f"worst gradient-norm deviation over {len(keys)} tensors: {worst:.4f}")
    for k in RC.GRAD_KEYS:
        a, b = grads[k].double().cpu().flatten(), torch.from_numpy(g["grad:" + k]).double().flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        print(f"  {k}: cosine {cos:.6f}")
        assert cos >= 0.999, k
    for (k, p), want in zip(sorted(net.named_parameters()), g["update_norms"]):
        got = float((p.detach() - before[k]).double().norm())
        assert abs(got - want) <= 0.02 * want + 1e-9, (k, got, want)


def test_raw_band_scaling_against_reference_run(cuda):
    """raw_image_folder.py:192-196 (run from the reference tree for the fixture) == licos_raw_dn_to_unit, bit for bit."""
    dn = torch.from_numpy(RC.dn_band().astype(np.int16)).to(cuda)
    full = ops.raw_dn_to_unit(dn, 4095, use_full_range=True).cpu().numpy()
    ubyte = ops.raw_dn_to_unit(dn, 4095, use_full_range=False).cpu().numpy()
    assert np.array_equal(full, GOLDEN["raw/full"][0])
    assert np.array_equal(ubyte, GOLDEN["raw/ubyte"][0])
