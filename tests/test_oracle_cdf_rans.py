"""Oracle self-checks for the integer path: C restatement vs its independent pure-Python twin, the committed
golden CDFs, and the rANS round trip incl. the bypass (escape) coding.  CPU only."""
import os

import numpy as np
import pytest

from oracle import cdf_rans as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_cdf_c_matches_python_twin_and_golden():
    z = np.load(os.path.join(GOLD, "cdf_cases.npz"))
    for i in range(4):
        pmf, gold = z[f"pmf{i}"], z[f"cdf{i}"]
        c = O.pmf_to_quantized_cdf(pmf.tolist(), 16)
        py = O.py_pmf_to_quantized_cdf(pmf.tolist(), 16)
        assert c == py == gold.tolist()
        assert c[0] == 0 and c[-1] == 1 << 16
        assert all(b > a for a, b in zip(c, c[1:])), "CDF must be strictly increasing"


def test_cdf_known_answers():
    # hand-checkable: uniform over 4 symbols
    assert O.pmf_to_quantized_cdf([0.25] * 4, 16) == [0, 16384, 32768, 49152, 65536]
    # a zero-mass symbol steals one count from the least-frequent donor with freq > 1 (first minimum wins)
    assert O.pmf_to_quantized_cdf([0.5, 0.0, 0.5], 16) == [0, 32767, 32768, 65536]
    assert O.py_pmf_to_quantized_cdf([0.5, 0.0, 0.5], 16) == [0, 32767, 32768, 65536]
    # GaussianConditional table row 0 (SURVEY.md 8c self-check 5)
    assert O.pmf_to_quantized_cdf([1e-9, 1.0 - 2e-9, 1e-9, 1e-9], 16) == [0, 1, 65534, 65535, 65536]


@pytest.mark.parametrize("bad", [[-0.1, 1.1], [float("nan"), 1.0], [float("inf"), 0.0], [0.0, 0.0]])
def test_cdf_rejects_bad_pmf(bad):
    with pytest.raises(ValueError):
        O.pmf_to_quantized_cdf(bad, 16)


def _tables(rng, n_cdfs=5, max_len=12):
    cdfs = np.zeros((n_cdfs, max_len + 2), dtype=np.int32)
    sizes = np.zeros(n_cdfs, dtype=np.int32)
    offsets = rng.integers(-6, 1, n_cdfs).astype(np.int32)
    for i in range(n_cdfs):
        ln = int(rng.integers(2, max_len + 1))
        p = rng.random(ln + 1).astype(np.float32)
        p /= p.sum()
        row = O.pmf_to_quantized_cdf(p.tolist(), 16)
        cdfs[i, : len(row)] = row
        sizes[i] = ln + 2
    return cdfs, sizes, offsets


def test_rans_roundtrip_with_escapes_and_twin():
    rng = np.random.default_rng(0)
    cdfs, sizes, offsets = _tables(rng)
    n = 4000
    idx = rng.integers(0, cdfs.shape[0], n).astype(np.int32)
    sym = rng.integers(-40, 40, n).astype(np.int32)  # far outside the tables -> bypass nibbles
    sym[:8] = [0, 1, -1, 300000, -300000, 2 ** 20, -(2 ** 20), 7]
    enc = O.encode_with_indexes(sym, idx, cdfs, sizes, offsets)
    assert len(enc) % 4 == 0
    dec = O.decode_with_indexes(enc, idx, cdfs, sizes, offsets)
    assert np.array_equal(dec, sym)
    # the pure-Python twin produces and reads the same bitstream
    m = 300
    enc_c = O.encode_with_indexes(sym[:m], idx[:m], cdfs, sizes, offsets)
    enc_py = O.py_encode_with_indexes(sym[:m].tolist(), idx[:m].tolist(), cdfs, sizes, offsets)
    assert enc_c == enc_py
    assert np.array_equal(O.py_decode_with_indexes(enc_c, idx[:m].tolist(), cdfs, sizes, offsets), sym[:m])


def test_rans_empty_and_single():
    rng = np.random.default_rng(1)
    cdfs, sizes, offsets = _tables(rng)
    enc = O.encode_with_indexes(np.zeros(0, np.int32), np.zeros(0, np.int32), cdfs, sizes, offsets)
    assert len(enc) == 8  # just the flushed 64-bit state
    one = O.encode_with_indexes(np.array([offsets[0]], np.int32), np.array([0], np.int32), cdfs, sizes, offsets)
    assert np.array_equal(O.decode_with_indexes(one, np.array([0], np.int32), cdfs, sizes, offsets), [offsets[0]])
