"""The cases that are run through the reference's own glue code (tests/ref_shim.py) to make tests/golden/reference_glue.npz,
and the seeded inputs both sides of the comparison rebuild.  Everything random is a function of a seed, so the GPU tests
(which have neither /root/reference nor the shim) reconstruct the same weights and inputs and compare the product with
the numbers the reference's code produced.

    eval cases   eval_utils.py:189-210 process_img (forward + compress, clamp, crop to the image size at :207),
                 :172-186 compute_bpp, :145-156 compute_psnr, :159-169 compute_msssim, on images whose sizes are NOT
                 multiples of 16, through models built by licos/model_utils.py:6-49 get_model
    train case   licos/train.py:148-212 train_one_batch: one step, optimizers from licos/utils.py:65-73
    raw bands    licos/raw_image_folder.py:183-196 _open_band_: 12-bit DN -> [0, 1] (-> 8-bit requantisation)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EVAL_CASES = {
    # name: (model, in_channels, quality, image (C, H, W), image seed)
    "rgb_q1": ("bmshj2018-factorized", 3, 1, (3, 180, 172), 101),       # 180 = 11.25 x 16, 172 = 10.75 x 16; > 160 for MS-SSIM
    "split_q1": ("bmshj2018-factorized", 1, 1, (1, 200, 264), 102),     # raw "split" single band, built from 12-bit DNs
    "merged_q1": ("bmshj2018-factorized", 13, 1, (13, 72, 100), 103),   # raw "merged" 13 bands (train.py:103-109)
    "relu_q1": ("bmshj2018-factorized-relu", 3, 1, (3, 90, 75), 104),   # odd width: not the TMA-aligned first-layer path
    # hyperprior: multiples of 64 (upstream's own forward raises on other sizes: h_s(h_a(y)) no longer matches y)
    "hyper_q1": ("bmshj2018-hyperprior", 3, 1, (3, 192, 256), 105),
    "hyper_q6": ("bmshj2018-hyperprior", 3, 6, (3, 128, 64), 106),      # N = 192, M = 320
}
TRAIN_CASE = ("bmshj2018-factorized", 3, 1, (4, 3, 64, 64), 201, 7)      # model, C, q, batch shape, data seed, noise seed
WEIGHT_SEED = 42
DN_SEED, DN_SHAPE = 301, (200, 264)


def synth_module():
    """licos_b200/synth.py loaded by path: the weight recipe without importing the package (and its CUDA library)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("licos_synth_standalone", os.path.join(ROOT, "licos_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def dn_band() -> np.ndarray:
    """Synthetic 12-bit digital numbers of one Sentinel-2 raw band (uint16 storage), incl. both extremes."""
    rng = np.random.default_rng(DN_SEED)
    dn = rng.integers(0, 4096, size=DN_SHAPE, dtype=np.uint16)
    dn[0, :4] = (0, 4095, 2047, 2048)
    return dn


def eval_image(name: str) -> torch.Tensor:
    model, c, q, shape, seed = EVAL_CASES[name]
    if name == "split_q1":  # the image a RawImageFolder band load returns (8-bit requantised, as the reference defaults)
        band = dn_band().astype(np.float64) / 4095
        return torch.from_numpy((np.clip(np.rint(band * 255.0), 0, 255).astype(np.uint8) / 255).astype(np.float32)).unsqueeze(0)
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g)


def build_weights(get_model, name: str):
    """Seeded default init through `get_model` (the reference's or a restatement's) + the synthetic conditioning."""
    model, c, q = name[:3] if isinstance(name, tuple) else EVAL_CASES[name][:3]
    torch.manual_seed(WEIGHT_SEED)
    net = get_model(model, False, c, q)
    synth_module().condition_weights(net)
    return net


def state_fingerprint(sd) -> np.ndarray:
    """float64 (sum, sum of squares) per floating-point state_dict entry, in key order."""
    rows = []
    for k in sorted(sd):
        v = sd[k]
        if v.is_floating_point() and v.numel():
            v = v.detach().double().cpu()
            rows.append((float(v.sum()), float((v * v).sum())))
    return np.asarray(rows, dtype=np.float64)


def run_eval_case(ref, name: str) -> dict:
    """ref: the namespace tests/ref_shim.reference() yields.  Returns plain numpy values."""
    net = build_weights(ref.model_utils.get_model, name)
    fingerprint = state_fingerprint(net.state_dict())  # before update(): parameters only, no coder tables
    net.eval()
    net.update()
    img = eval_image(name)
    out_net, reconstructed, diff, nbytes = ref.eval_utils.process_img(img, net)
    x_hat = out_net["x_hat"]
    res = {
        "fingerprint": fingerprint,
        "bytes": np.int64(nbytes),
        "bpp": np.float64(ref.eval_utils.compute_bpp(out_net)),
        "psnr": np.float64(ref.eval_utils.compute_psnr(img.unsqueeze(0), x_hat.cpu())),
        "x_hat_shape": np.asarray(x_hat.shape, dtype=np.int64),
        "x_hat_mean": x_hat.double().mean(dim=(0, 2, 3)).cpu().numpy(),
        "x_hat_lowres": torch.nn.functional.adaptive_avg_pool2d(x_hat.float().cpu(), (8, 8)).numpy(),
        "diff_mean": np.float64(diff.double().mean()),
        "recon_shape": np.asarray(reconstructed.shape, dtype=np.int64),
        "lik_shapes": np.asarray([list(v.shape) for v in out_net["likelihoods"].values()], dtype=np.int64),
    }
    if min(img.shape[1:]) > 160:
        res["msssim"] = np.float64(ref.eval_utils.compute_msssim(img.unsqueeze(0), x_hat.cpu()))
    return res


def train_inputs():
    model, c, q, shape, data_seed, noise_seed = TRAIN_CASE
    g = torch.Generator().manual_seed(data_seed)
    return torch.rand(*shape, generator=g)


def train_noise(latent_shape) -> torch.Tensor:
    """The noise tensor the oracle's EntropyBottleneck draws as the FIRST random draw of the step after
    torch.manual_seed(noise seed): shape (C, 1, B*h*w) in the permuted order, returned as (B, C, h, w)."""
    B, C, h, w = latent_shape
    torch.manual_seed(TRAIN_CASE[5])
    nz = torch.empty(C, 1, B * h * w).uniform_(-0.5, 0.5)
    return nz.reshape(C, B, h, w).permute(1, 0, 2, 3).contiguous()


class RecordingCriterion(torch.nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.inner, self.last = inner, None

    def forward(self, out, target):
        self.last = self.inner(out, target)
        return self.last


GRAD_KEYS = ("g_a.0.bias", "g_a.1.beta", "g_a.6.bias", "g_s.0.bias", "g_s.5.beta", "g_s.6.weight", "g_s.6.bias",
             "entropy_bottleneck._matrix0", "entropy_bottleneck._bias2", "entropy_bottleneck._factor1",
             "entropy_bottleneck.quantiles")


def run_train_case(ref, RateDistortionLoss) -> dict:
    """One call of the reference's train_one_batch (train.py:148-212) on a seeded batch."""
    model, c, q, shape, data_seed, noise_seed = TRAIN_CASE
    net = build_weights(ref.model_utils.get_model, (model, c, q))
    cfg = types.SimpleNamespace(learning_rate=1e-4, aux_learning_rate=1e-3)
    optimizer, aux_optimizer = ref.utils.configure_optimizers(net, cfg)
    crit = RecordingCriterion(RateDistortionLoss(lmbda=1e-2))
    d = train_inputs()
    before = {k: v.detach().clone() for k, v in net.named_parameters()}
    torch.manual_seed(noise_seed)
    it = ref.train.train_one_batch(0, net, crit, [d], iter([d]), optimizer, aux_optimizer, 1, 1.0)
    assert it is not None and net.training
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    res = {
        "loss": np.float64(crit.last["loss"].item()), "mse_loss": np.float64(crit.last["mse_loss"].item()),
        "bpp_loss": np.float64(crit.last["bpp_loss"].item()),
        "aux_loss": np.float64(net.aux_loss().item()),  # after the aux step
        "grad_keys": np.asarray(sorted(grads)),
        "grad_norms": np.asarray([float(grads[k].double().norm()) for k in sorted(grads)], dtype=np.float64),
        "update_norms": np.asarray([float((p.detach() - before[k]).double().norm()) for k, p in sorted(net.named_parameters())],
                                   dtype=np.float64),
        "aux_param_names": np.asarray(sorted(k for k in grads if k.endswith(".quantiles"))),
        "n_net_params": np.int64(sum(len(g["params"]) for g in optimizer.param_groups)),
        "n_aux_params": np.int64(sum(len(g["params"]) for g in aux_optimizer.param_groups)),
    }
    for k in GRAD_KEYS:
        res["grad:" + k] = grads[k].numpy()
    return res


def run_raw_band_case(ref, tmpdir: str) -> dict:
    """raw_image_folder.py:183-196 on a synthetic band (rasterio.open is the stand-in that reads an .npy file)."""
    path = os.path.join(tmpdir, "band.npy")
    np.save(path, dn_band())
    out = {}
    for full in (True, False):
        fake_self = types.SimpleNamespace(use_full_range=full)
        t = ref.raw_image_folder.RawImageFolder._open_band_(fake_self, path)
        out["full" if full else "ubyte"] = t.numpy()
    out["dn_max"] = np.int64(ref.raw_utils.DN_MAX)
    return out
