"""Generates tests/golden/*.npz with the CPU oracle (oracle/compressai_ref.py).

PARITY UNPINNED: the reference holds no golden vectors for this path and `compressai` is not installable here
(SURVEY.md section 8c), so these fixtures pin the ORACLE's outputs on seeded inputs -- they catch drift of the
oracle / the weight recipe and let the GPU tests check the CUDA path without importing the oracle's model.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from licos_b200 import synth  # noqa: E402  (weight recipe shared with the product tests / bench)
from oracle import compressai_ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build(name, quality, in_ch, seed=42):
    torch.manual_seed(seed)
    net = R.image_models[name](quality=quality) if in_ch == 3 else R.get_model(name, False, in_ch, quality)
    synth.condition_weights(net)
    net.eval()
    net.update(force=True)
    return net


def factorized_case(in_ch, tag):
    net = build("bmshj2018-factorized", 1, in_ch)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, in_ch, 64, 64, generator=g)
    with torch.no_grad():
        y = net.g_a(x)
        out = net(x)
        comp = net.compress(x)
        eb = net.entropy_bottleneck
        sym = eb.quantize(y, "symbols", eb._get_medians().detach().reshape(1, -1, 1, 1))
        y_hat, lik = eb(y)
        ntrain = torch.rand(y.shape, generator=g) - 0.5
        y_noisy, lik_noisy = eb(y, training=True, noise=ntrain)
    num_pixels = x.size(0) * x.size(2) * x.size(3)
    np.savez_compressed(
        os.path.join(HERE, f"factorized_{tag}.npz"),
        x=x.numpy(), y=y.numpy(), symbols=sym.numpy(), y_hat=y_hat.numpy(), lik=lik.numpy(),
        noise=ntrain.numpy(), y_noisy=y_noisy.numpy(), lik_noisy=lik_noisy.numpy(),
        x_hat=out["x_hat"].numpy(),
        bpp=np.float64(torch.log(out["likelihoods"]["y"]).sum().item() / (-np.log(2) * num_pixels)),
        mse=np.float64(torch.mean((out["x_hat"] - x) ** 2).item()),
        quantized_cdf=eb._quantized_cdf.numpy(), cdf_length=eb._cdf_length.numpy(), offset=eb._offset.numpy(),
        string_lengths=np.array([len(s) for s in comp["strings"][0]]),
        string_sha=np.array([hashlib.sha256(s).hexdigest() for s in comp["strings"][0]]),
        string0=np.frombuffer(comp["strings"][0][0], dtype=np.uint8),
        nonzero_fraction=np.float64((sym != 0).float().mean().item()),
    )
    print(tag, "nonzero symbols", (sym != 0).float().mean().item(), "|sym| max", sym.abs().max().item(),
          "bytes", [len(s) for s in comp["strings"][0]])


def hyperprior_case():
    net = build("bmshj2018-hyperprior", 1, 3)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(1, 3, 128, 128, generator=g)
    with torch.no_grad():
        y = net.g_a(x)
        z = net.h_a(torch.abs(y))
        z_hat, z_lik = net.entropy_bottleneck(z)
        scales = net.h_s(z_hat)
        y_hat, y_lik = net.gaussian_conditional(y, scales)
        idx = net.gaussian_conditional.build_indexes(scales)
        out = net(x)
        comp = net.compress(x)
    gc = net.gaussian_conditional
    np.savez_compressed(
        os.path.join(HERE, "hyperprior_rgb.npz"),
        x=x.numpy(), y=y.numpy(), z=z.numpy(), z_hat=z_hat.numpy(), z_lik=z_lik.numpy(), scales=scales.numpy(),
        y_hat=y_hat.numpy(), y_lik=y_lik.numpy(), indexes=idx.numpy(), x_hat=out["x_hat"].numpy(),
        scale_table=gc.scale_table.numpy(), gc_cdf_sha=np.array(sha(gc._quantized_cdf.numpy())),
        gc_cdf_row0=gc._quantized_cdf[0, :5].numpy(), gc_cdf_length=gc._cdf_length.numpy(),
        gc_offset=gc._offset.numpy(),
        y_string_lengths=np.array([len(s) for s in comp["strings"][0]]),
        z_string_lengths=np.array([len(s) for s in comp["strings"][1]]),
        y_string_sha=np.array([hashlib.sha256(s).hexdigest() for s in comp["strings"][0]]),
        z_string_sha=np.array([hashlib.sha256(s).hexdigest() for s in comp["strings"][1]]),
    )
    print("hyperprior: scales range", scales.min().item(), scales.max().item(), "idx range",
          idx.min().item(), idx.max().item(), "bytes", [len(s) for s in comp["strings"][0]],
          [len(s) for s in comp["strings"][1]])


def cdf_case():
    rng = np.random.default_rng(3)
    pmfs, cdfs = [], []
    for n in (2, 5, 33, 257):
        p = rng.random(n).astype(np.float32) ** 4
        p[rng.integers(0, n, size=max(1, n // 4))] = 0  # forces frequency stealing
        p = p / p.sum()
        pmfs.append(p)
        cdfs.append(np.array(R.cdf_rans.pmf_to_quantized_cdf(p.tolist(), 16), dtype=np.uint32))
    np.savez_compressed(os.path.join(HERE, "cdf_cases.npz"), **{f"pmf{i}": p for i, p in enumerate(pmfs)},
                        **{f"cdf{i}": c for i, c in enumerate(cdfs)})


if __name__ == "__main__":
    factorized_case(3, "rgb")
    factorized_case(1, "split")
    factorized_case(13, "merged")
    hyperprior_case()
    cdf_case()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
