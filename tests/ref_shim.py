"""Runs the reference's OWN glue files (/root/reference/licos/*.py, /root/reference/eval_utils.py) unmodified, with the
third-party packages they import replaced by stand-ins:

    compressai.{zoo, entropy_models, losses, optimizers, datasets, registry}  ->  ``oracle.compressai_ref`` (CPU checker)
                                                                               or ``licos_b200`` (the product)
    pytorch_msssim.ms_ssim                                                    ->  ``oracle.msssim_ref`` / ``licos_b200.ops``
    dotmap.DotMap, rasterio, skimage.img_as_ubyte                             ->  minimal stand-ins below

Nothing is copied: the reference files are imported from where they lie and only exist in the build container
(``/root/reference`` does not travel to the GPU box), so everything that uses this module is a ``not gpu`` test or the
fixture generator ``tests/golden/make_reference_golden.py`` and skips when the tree is absent.

The reference files this executes:
    licos/model_utils.py:6-49      get_model (zoo lookup + bottleneck / first conv / last deconv surgery)
    licos/train.py:148-212         train_one_batch (the training step body)
    licos/train.py:262-303         test_epoch
    licos/utils.py:65-73           configure_optimizers (net / aux split)
    licos/federation_utils.py:8-85 update_central_model (file-based weighted merge)
    licos/raw_image_folder.py:183-196  _open_band_ (DN -> [0,1] -> optional 8-bit requantisation)
    eval_utils.py:145-210          compute_psnr, compute_msssim, compute_bpp, process_img
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REFERENCE = "/root/reference"
_REF_MODULES = ("model_utils", "utils", "train", "federation_utils", "eval_utils", "raw_image_folder", "raw_utils")
_SHIMMED = ("compressai", "compressai.zoo", "compressai.entropy_models", "compressai.losses", "compressai.optimizers",
            "compressai.datasets", "compressai.registry", "compressai.layers", "compressai.models", "pytorch_msssim",
            "dotmap", "rasterio", "skimage")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE, "licos", "model_utils.py"))


class DotMap(dict):
    """dotmap.DotMap as far as the reference uses it: attribute access on a dict (cfg.save_path, cfg.seed, ...)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _img_as_ubyte(a):
    """skimage.img_as_ubyte for float input in [0, 1] (its documented conversion: scale by 255, round to nearest even,
    clip) -- the only use the reference makes of it (raw_image_folder.py:195)."""
    import numpy as np

    a = np.asarray(a)
    if a.dtype.kind != "f":
        raise TypeError("stand-in covers float input only")
    if a.min() < -1.0 or a.max() > 1.0:
        raise ValueError("Images of type float must be between -1 and 1.")
    return np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8)


class _FakeRaster:
    def __init__(self, path):
        import numpy as np

        self._a = np.load(path)

    def read(self, band):
        return self._a

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _backend_modules(backend: str):
    if backend == "oracle":
        from oracle import compressai_ref as B
        from oracle import msssim_ref

        ms_ssim = msssim_ref.ms_ssim
    elif backend == "licos_b200":
        import licos_b200 as B
        from licos_b200 import ops

        ms_ssim = ops.ms_ssim
    else:
        raise ValueError(backend)
    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        mods[name] = m
        return m

    pkg = mod("compressai")
    pkg.__path__ = []  # a package, so `from compressai.zoo import ...` resolves through sys.modules
    mod("compressai.zoo", image_models=B.image_models)
    mod("compressai.entropy_models", EntropyBottleneck=B.EntropyBottleneck, GaussianConditional=B.GaussianConditional,
        EntropyModel=B.EntropyModel)
    mod("compressai.losses", RateDistortionLoss=B.RateDistortionLoss)
    mod("compressai.optimizers", net_aux_optimizer=B.net_aux_optimizer)
    mod("compressai.layers", GDN=B.GDN)
    mod("compressai.models", FactorizedPrior=B.FactorizedPrior, ScaleHyperprior=B.ScaleHyperprior,
        CompressionModel=B.CompressionModel)
    mod("compressai.datasets", ImageFolder=type("ImageFolder", (), {}))
    mod("compressai.registry", register_dataset=lambda name: (lambda cls: cls))
    mod("pytorch_msssim", ms_ssim=ms_ssim)
    mod("dotmap", DotMap=DotMap)
    mod("rasterio", open=_FakeRaster)
    mod("skimage", img_as_ubyte=_img_as_ubyte)
    return mods


@contextlib.contextmanager
def reference(backend: str):
    """``with reference("oracle") as ref: ref.model_utils.get_model(...)`` -- the reference's modules, freshly executed
    against the chosen backend; sys.modules / sys.path are restored on exit."""
    if not available():
        raise FileNotFoundError(REFERENCE)
    saved = {k: sys.modules.get(k) for k in _SHIMMED + _REF_MODULES}
    saved_path = list(sys.path)
    try:
        for k in _SHIMMED + _REF_MODULES:
            sys.modules.pop(k, None)
        sys.modules.update(_backend_modules(backend))
        sys.path[:0] = [os.path.join(REFERENCE, "licos"), REFERENCE]
        ns = types.SimpleNamespace(backend=backend)
        for name in _REF_MODULES:
            setattr(ns, name, importlib.import_module(name))
        for name in _REF_MODULES:  # the files that were executed really are the reference's
            assert getattr(ns, name).__file__.startswith(REFERENCE + os.sep), getattr(ns, name).__file__
        yield ns
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
