"""The C-ABI shared library loads and exports every symbol include/licos_b200.h declares; the host-only entry
points (CDF, rANS) are exercised against the oracle.  No GPU compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from licos_b200 import _lib, ops
from oracle import cdf_rans as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _declared():
    src = open(os.path.join(ROOT, "include", "licos_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(licos_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in licos_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names) | {"licos_set_last_cuda_error"}
    assert lib.licos_abi_version() == _lib.ABI_VERSION == 3


def test_struct_layouts_match_header():
    # licos_conv_args: 9 ints, 6 pointers, pointer + int64, 2 ints
    assert ctypes.sizeof(_lib.ConvArgs) == 9 * 4 + 4 + 7 * 8 + 8 + 2 * 4 + 8  # ..., sm_count, int_max, pre_act
    assert _lib.ConvArgs.in_.offset == 40 and _lib.ConvArgs.workspace_bytes.offset == 96
    assert _lib.EbParams.packed.offset == 48 and ctypes.sizeof(_lib.EbParams) == 72


def test_error_strings_and_status_mapping():
    assert _lib.lib.licos_strerror(0) == b"ok"
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(-3, "x")
    with pytest.raises(_lib.LicosError):
        _lib.check(-4, "x")
    assert _lib.lib.licos_packed_weight_bytes(_lib.CONV_5X5_S2, 128, 128, _lib.LAYOUT_NHWC_BF16) == 25 * 128 * 128 * 2
    assert _lib.lib.licos_packed_weight_bytes(_lib.CONV_5X5_S2, 320, 192, _lib.LAYOUT_NHWC_BF16) == 25 * 320 * 192 * 2
    assert _lib.lib.licos_packed_weight_bytes(_lib.CONV_5X5_S2, 128, 3, _lib.LAYOUT_NCHW_F32) == 128 * 128 * 2
    assert _lib.lib.licos_packed_weight_bytes(_lib.DECONV_5X5_S2, 3, 128, _lib.LAYOUT_NHWC_BF16) == 3 * 48 * 128 * 2
    assert _lib.lib.licos_packed_weight_bytes(_lib.DECONV_5X5_S2, 13, 128, _lib.LAYOUT_NHWC_BF16) == 25 * 16 * 128 * 2


def test_pmf_to_quantized_cdf_matches_oracle_and_golden():
    z = np.load(os.path.join(GOLD, "cdf_cases.npz"))
    for i in range(4):
        assert ops.pmf_to_quantized_cdf(z[f"pmf{i}"]).tolist() == z[f"cdf{i}"].tolist()
    rng = np.random.default_rng(5)
    for _ in range(200):
        n = int(rng.integers(1, 80))
        p = rng.random(n).astype(np.float32) ** int(rng.integers(1, 9))
        p[rng.random(n) < 0.3] = 0
        if p.sum() == 0:
            p[0] = 1
        p /= p.sum()
        assert ops.pmf_to_quantized_cdf(p).tolist() == O.pmf_to_quantized_cdf(p.tolist(), 16)
    for bad in ([-0.1, 1.1], [float("nan"), 1.0], [float("inf"), 0.0], [0.0, 0.0]):
        with pytest.raises(ValueError):
            ops.pmf_to_quantized_cdf(bad)


def _tables(rng, n_cdfs=7, max_len=20):
    cdfs = np.zeros((n_cdfs, max_len + 2), dtype=np.int32)
    sizes = np.zeros(n_cdfs, dtype=np.int32)
    offsets = rng.integers(-9, 1, n_cdfs).astype(np.int32)
    for i in range(n_cdfs):
        ln = int(rng.integers(2, max_len + 1))
        p = rng.random(ln + 1).astype(np.float32)
        p /= p.sum()
        row = O.pmf_to_quantized_cdf(p.tolist(), 16)
        cdfs[i, : len(row)] = row
        sizes[i] = ln + 2
    return cdfs, sizes, offsets


def test_rans_bitstream_identical_to_oracle():
    rng = np.random.default_rng(2)
    cdfs, sizes, offsets = _tables(rng)
    for n in (0, 1, 17, 5000):
        idx = rng.integers(0, cdfs.shape[0], n).astype(np.int32)
        sym = np.rint(rng.normal(0, 6, n)).astype(np.int32)
        if n > 10:
            sym[:6] = [10 ** 6, -(10 ** 6), 255, -256, 0, 16]
        enc = ops.rans_encode(sym, idx, cdfs, sizes, offsets)
        assert enc == O.encode_with_indexes(sym, idx, cdfs, sizes, offsets)
        assert np.array_equal(ops.rans_decode(enc, idx, cdfs, sizes, offsets), sym)
        assert np.array_equal(O.decode_with_indexes(enc, idx, cdfs, sizes, offsets), sym)


def test_rans_batch_threads_and_shared_index_plane():
    rng = np.random.default_rng(3)
    cdfs, sizes, offsets = _tables(rng)
    B, n = 9, 700
    idx = rng.integers(0, cdfs.shape[0], n).astype(np.int32)
    sym = np.rint(rng.normal(0, 5, (B, n))).astype(np.int32)
    one = ops.rans_encode_batch(sym, idx, cdfs, sizes, offsets, threads=1)
    many = ops.rans_encode_batch(sym, np.tile(idx, (B, 1)), cdfs, sizes, offsets, threads=4)
    assert one == many == [O.encode_with_indexes(sym[i], idx, cdfs, sizes, offsets) for i in range(B)]
    assert np.array_equal(ops.rans_decode_batch(one, idx, n, cdfs, sizes, offsets, threads=3), sym)
    with pytest.raises(ValueError):
        ops.rans_encode(sym[0], np.full(n, 99, np.int32), cdfs, sizes, offsets)  # index out of range
