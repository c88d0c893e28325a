"""Generates tests/golden/reference_glue.npz by executing the REFERENCE'S OWN FILES (/root/reference/licos/model_utils.py,
train.py, utils.py, raw_image_folder.py and /root/reference/eval_utils.py, unmodified, imported from where they lie) with
`compressai` / `pytorch_msssim` replaced by the CPU oracle (tests/ref_shim.py).  The cases are listed in
tests/reference_cases.py.  Runs only in the build container (the reference tree does not travel to the GPU box); the
fixture does, and tests/test_gpu_reference_golden.py compares the CUDA path with it.

What this pins, and what it cannot: the glue semantics (surgery, step order, crop / clamp, bpp / PSNR definitions, the
optimizer split, DN scaling) are the reference's own code; the arithmetic under the model API is still the oracle's
restatement of CompressAI (not installable here) -- DESIGN.md section 2.

    python tests/golden/make_reference_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.path.insert(0, os.path.dirname(TESTS))

import ref_shim  # noqa: E402
import reference_cases as RC  # noqa: E402


def generate() -> dict:
    from oracle import compressai_ref as R

    out = {}
    with ref_shim.reference("oracle") as ref:
        for name in RC.EVAL_CASES:
            for k, v in RC.run_eval_case(ref, name).items():
                out[f"eval/{name}/{k}"] = v
        for k, v in RC.run_train_case(ref, R.RateDistortionLoss).items():
            out[f"train/{k}"] = v
        with tempfile.TemporaryDirectory() as tmp:
            for k, v in RC.run_raw_band_case(ref, tmp).items():
                out[f"raw/{k}"] = v
    return out


if __name__ == "__main__":
    import torch

    torch.set_num_threads(1)  # a fixed reduction order: the fixture does not depend on the host's core count
    data = generate()
    path = os.path.join(HERE, "reference_glue.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "entries")
    for k in sorted(data):
        if data[k].size == 1:
            print(f"  {k} = {data[k]}")
