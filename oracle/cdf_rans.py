"""ctypes binding of oracle/cdf_rans.c plus a pure-Python twin for small cases.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.
The Python twin restates the same upstream algorithms independently of the C
file so the two can be checked against each other (tests/test_oracle_cdf_rans.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_cdf_rans.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cdf_rans.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_cdf_rans.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        lib.oracle_pmf_to_quantized_cdf.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.oracle_rans_encode_with_indexes.restype = ctypes.c_long
        lib.oracle_rans_encode_with_indexes.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        lib.oracle_rans_decode_with_indexes.restype = ctypes.c_int
        lib.oracle_rans_decode_with_indexes.argtypes = [
            ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def pmf_to_quantized_cdf(pmf, precision: int = 16):
    lib = _load()
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    cdf = np.zeros(p.size + 1, dtype=np.uint32)
    rc = lib.oracle_pmf_to_quantized_cdf(p.ctypes.data, p.size, precision, cdf.ctypes.data)
    if rc == -1:
        raise ValueError("Invalid `pmf`, non-finite or negative element found")
    if rc == -2:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability")
    if rc != 0:
        raise ValueError("pmf_to_quantized_cdf: no frequency left to steal")
    return cdf.tolist()


def encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes:
    lib = _load()
    symbols, indexes = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    cdfs, cdfs_sizes, offsets = _i32(cdfs), _i32(cdfs_sizes), _i32(offsets)
    assert cdfs.ndim == 2
    cap = 4 * (symbols.size * 12 + 16)
    out = np.empty(cap, dtype=np.uint8)
    n = lib.oracle_rans_encode_with_indexes(
        symbols.ctypes.data, indexes.ctypes.data, symbols.size, cdfs.ctypes.data,
        cdfs.shape[0], cdfs.shape[1], cdfs_sizes.ctypes.data, offsets.ctypes.data,
        out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(f"oracle rans encode failed ({n})")
    return out[:n].tobytes()


def decode_with_indexes(encoded: bytes, indexes, cdfs, cdfs_sizes, offsets) -> np.ndarray:
    lib = _load()
    indexes = _i32(indexes).reshape(-1)
    cdfs, cdfs_sizes, offsets = _i32(cdfs), _i32(cdfs_sizes), _i32(offsets)
    enc = np.frombuffer(encoded, dtype=np.uint8)
    out = np.empty(indexes.size, dtype=np.int32)
    rc = lib.oracle_rans_decode_with_indexes(
        enc.ctypes.data, enc.size, indexes.ctypes.data, indexes.size, cdfs.ctypes.data,
        cdfs.shape[0], cdfs.shape[1], cdfs_sizes.ctypes.data, offsets.ctypes.data,
        out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle rans decode failed ({rc})")
    return out


# ---------------------------------------------------------------------------
# pure-Python twin (small cases only)
# ---------------------------------------------------------------------------

def py_pmf_to_quantized_cdf(pmf, precision: int = 16):
    f32 = np.float32
    vals = [f32(p) for p in pmf]
    for p in vals:
        if p < 0 or not np.isfinite(p):
            raise ValueError("Invalid `pmf`, non-finite or negative element found")
    scale = f32(1 << precision)
    cdf = [0]
    for p in vals:
        x = float(f32(p * scale))
        # C roundf: half away from zero
        cdf.append(int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5)))
    total = sum(cdf)
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability")
    cdf = [((1 << precision) * c) // total for c in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    n = len(cdf) - 1
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            best_freq, best = None, -1
            for j in range(n):
                fr = cdf[j + 1] - cdf[j]
                if fr > 1 and (best_freq is None or fr < best_freq):
                    best_freq, best = fr, j
            assert best != -1
            if best < i:
                for j in range(best + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best + 1):
                    cdf[j] += 1
    return cdf


_L = 1 << 31
_M32 = (1 << 32) - 1


def py_encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes:
    syms = []
    for s, ci in zip(symbols, indexes):
        cdf = cdfs[ci]
        max_value = int(cdfs_sizes[ci]) - 2
        value = int(s) - int(offsets[ci])
        raw = 0
        if value < 0:
            raw, value = -2 * value - 1, max_value
        elif value >= max_value:
            raw, value = 2 * (value - max_value), max_value
        syms.append((int(cdf[value]), int(cdf[value + 1] - cdf[value]), False))
        if value == max_value:
            nb = 0
            while (raw >> (nb * 4)) != 0:
                nb += 1
            v = nb
            while v >= 15:
                syms.append((15, 16, True))
                v -= 15
            syms.append((v, v + 1, True))
            for j in range(nb):
                v = (raw >> (j * 4)) & 15
                syms.append((v, v + 1, True))
    words = []
    x = _L
    for start, rng, bypass in reversed(syms):
        if not bypass:
            x_max = ((_L >> 16) << 32) * rng
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = ((x // rng) << 16) + (x % rng) + start
        else:
            freq = 1 << (16 - 4)
            x_max = ((_L >> 16) << 32) * freq
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = (x << 4) | start
    words.append((x >> 32) & _M32)
    words.append(x & _M32)
    words.reverse()
    return np.asarray(words, dtype=np.uint32).tobytes()


def py_decode_with_indexes(encoded: bytes, indexes, cdfs, cdfs_sizes, offsets):
    words = list(np.frombuffer(encoded, dtype=np.uint32)) + [0] * 4
    words = [int(w) for w in words]
    x = words[0] | (words[1] << 32)
    pos = 2
    out = []

    def get_bits(n):
        nonlocal x, pos
        val = x & ((1 << n) - 1)
        x >>= n
        if x < _L:
            x = (x << 32) | words[pos]
            pos += 1
        return val

    for ci in indexes:
        cdf = cdfs[ci]
        max_value = int(cdfs_sizes[ci]) - 2
        cum = x & 0xFFFF
        k = 0
        while k < cdfs_sizes[ci] and not (int(cdf[k]) > cum):
            k += 1
        s = k - 1
        start, freq = int(cdf[s]), int(cdf[s + 1] - cdf[s])
        x = freq * (x >> 16) + (x & 0xFFFF) - start
        if x < _L:
            x = (x << 32) | words[pos]
            pos += 1
        value = s
        if value == max_value:
            v = get_bits(4)
            nb = v
            while v == 15:
                v = get_bits(4)
                nb += v
            raw = 0
            for j in range(nb):
                raw |= get_bits(4) << (j * 4)
            value = raw >> 1
            if raw & 1:
                value = -value - 1
            else:
                value += max_value
        out.append(value + int(offsets[ci]))
    return np.asarray(out, dtype=np.int32)
