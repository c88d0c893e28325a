// Device-resident front end of compress() (SURVEY.md section 8f rows N1 stretch / N2): the rANS ENCODER on the GPU,
// bitstream-identical to csrc/host_codec.cpp (and therefore to compressai.ans: 64-bit state, 32-bit renormalisation
// words, 16-bit precision, 4-bit bypass nibbles with the base-15 count prefix, symbols coded in reverse), so that the
// symbols never leave the device: only the finished byte strings are copied to the host.
//
// A rANS stream is sequential in its state, and the CompressAI format has ONE stream per image, so the parallelism is
// across the images of a batch: one thread per image, one warp per block so the streams spread over up to `batch / 32`
// SMs.  The exact 64-by-16-bit division the coder needs (x / range, x % range) is one mul.hi.u64 + a correction against
// a per-table-entry reciprocal (floor((2^64 - 1) / range), rans_rcp_kernel); symbols and their table entries are fetched 16
// at a time ahead of the sequential state updates, so the walk pays one DRAM latency per 16 symbols.  Words are pushed from the end of a per-image scratch row
// (the stream is written back to front); a second kernel packs the rows into one contiguous buffer.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/licos_b200.h"
#include <type_traits>

#include "common.cuh"

namespace licos {

constexpr uint32_t kRansPrecision = 16, kRansBypassBits = 4, kRansBypassMax = 15;
constexpr uint64_t kRansLow = 1ull << 31;

// The coder state is kept as two 32-bit halves: after the renormalisation test the update needs x / range and
// x % range (64 by 16 bit).  With rcp = floor((2^64 - 1) / range): q0 = mulhi64(x, rcp) is q or q - 1, and the
// remainder is below 2^17, so it is computed in 32-bit arithmetic.  The new low word is (q_lo << 16) | (r + start):
// r + start < 2^16, no carry.  Everything is straight-line, predicated code: a lone warp per SM pays the full latency of
// every dependent instruction and every convergence barrier, so the loop body is kept short and branch-free.
struct RansState {
    uint32_t lo, hi;
    uint32_t* wp;  // words are pushed downwards from the end of this image's scratch row: *--wp
};

__device__ __forceinline__ void rans_put_symbol(RansState& s, uint32_t start, uint32_t range, uint64_t rcp) {
    const bool ren = s.hi >= (range << 15);  // x >= ((L >> 16) << 32) * range = range << 47
    if (ren) {
        *--s.wp = s.lo;
        s.lo = s.hi;
        s.hi = 0;
    }
    const uint64_t x = ((uint64_t)s.hi << 32) | s.lo;
    uint64_t q = __umul64hi(x, rcp);
    uint32_t r = s.lo - (uint32_t)q * range;  // the true remainder (plus at most one range) fits 32 bits
    if (r >= range) { r -= range; ++q; }
    s.hi = (uint32_t)(q >> 16);               // x' = (q << 16) + r + start
    s.lo = ((uint32_t)q << 16) | (r + start);
}

__device__ __forceinline__ void rans_put_nibble(RansState& s, uint32_t nib) {
    if (s.hi >= (1u << 27)) {  // x >= ((L >> 16) << 32) * 2^(16 - 4) = 2^59
        *--s.wp = s.lo;
        s.lo = s.hi;
        s.hi = 0;
    }
    s.hi = (s.hi << kRansBypassBits) | (s.lo >> (32 - kRansBypassBits));
    s.lo = (s.lo << kRansBypassBits) | nib;
}

// Bypass coding of an out-of-support symbol.  Forward order is [symbol][count prefix: 15, 15, .., rest][nibbles low to
// high]; coding runs in reverse.  Out of line: the hot loop is unrolled 16 times and must stay small (a lone warp per
// SM is at the mercy of the instruction cache).
__device__ __noinline__ void rans_put_escape(RansState& st, uint32_t raw) {
    uint32_t nibbles = 0;
    while (nibbles < 8 && (raw >> (nibbles * kRansBypassBits)) != 0) ++nibbles;  // (a shift by 32 is undefined)
    for (uint32_t j = nibbles; j-- > 0;) rans_put_nibble(st, (raw >> (j * kRansBypassBits)) & kRansBypassMax);
    rans_put_nibble(st, nibbles % kRansBypassMax);
    for (uint32_t c15 = nibbles / kRansBypassMax; c15 > 0; --c15) rans_put_nibble(st, kRansBypassMax);
}

// reciprocal of every table frequency, once per call: rcp[ci][v] = floor((2^64 - 1) / (cdf[v+1] - cdf[v]))
__global__ void rans_rcp_kernel(const int32_t* __restrict__ cdfs, int n_cdfs, int stride, const int32_t* __restrict__ sizes,
                                uint64_t* __restrict__ rcp) {
    const int64_t total = (int64_t)n_cdfs * stride;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(e / stride), v = (int)(e % stride);
        uint64_t r = 0;
        if (v + 1 < sizes[ci]) {
            const uint32_t f = (uint32_t)(cdfs[e + 1] - cdfs[e]);
            if (f > 0) r = 0xFFFFFFFFFFFFFFFFull / f;
        }
        rcp[e] = r;
    }
}

// symbols [batch][n] int32; indexes: explicit [batch][n] (or [n] shared: idx_stride 0), or NULL -> index = i / n_spatial
__global__ void __launch_bounds__(32) rans_encode_kernel(const int32_t* __restrict__ symbols, const int32_t* __restrict__ indexes,
                                                         int64_t idx_stride, int batch, int64_t n, int64_t n_spatial,
                                                         const int32_t* __restrict__ cdfs, int n_cdfs, int stride,
                                                         const int32_t* __restrict__ sizes, const int32_t* __restrict__ offsets,
                                                         const uint64_t* __restrict__ rcp, uint32_t* __restrict__ work,
                                                         int64_t cap_words, int32_t* __restrict__ lengths) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= batch) return;
    const int32_t* sym = symbols + (size_t)b * n;
    const int32_t* idx = indexes ? indexes + (size_t)b * idx_stride : nullptr;
    uint32_t* const row = work + (size_t)b * cap_words;
    RansState st{(uint32_t)kRansLow, 0u, row + cap_words};
    bool bad = false;
    // Walk the symbols backwards in chunks of 16.  Phase A (independent of the coder state): fetch the symbols and look
    // up (start, range, reciprocal) for all 16 -- the loads overlap, one DRAM latency per chunk.  Phase B: the
    // sequential state updates, from registers.  Out-of-support symbols (bypass nibbles) are re-derived in phase B.
    constexpr int kChunk = 16;
    const bool vec = ((((uintptr_t)sym) | (uintptr_t)(n * 4)) & 15) == 0;  // every full chunk start is 16-byte aligned
    int64_t i = n;
    // channel index of the chunk start without a division per symbol (index == NULL: symbol i belongs to table i / n_spatial)
    int64_t first = n - (n % kChunk ? n % kChunk : (n ? kChunk : 0));
    int32_t ci_b = (int32_t)(first / n_spatial);
    int64_t rem_b = first % n_spatial;
    while (i > 0) {
        const int cnt = (int)(i % kChunk ? i % kChunk : kChunk);  // the ragged chunk comes first: the rest are full and aligned
        const int64_t base = i - cnt;
        if (st.wp - row < 4 * kChunk * 3) { bad = true; break; }  // scratch row nearly full: give up, the host coder takes over
        int32_t sv[kChunk];
        if (vec && cnt == kChunk) {
#pragma unroll
            for (int q = 0; q < kChunk / 4; ++q) {
                const int4 t = __ldg(reinterpret_cast<const int4*>(sym + base) + q);
                sv[4 * q] = t.x; sv[4 * q + 1] = t.y; sv[4 * q + 2] = t.z; sv[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kChunk; ++k) sv[k] = k < cnt ? __ldg(sym + base + k) : 0;
        }
        uint32_t sta[kChunk], rg[kChunk];  // start; range (0 = invalid) | bit 31 = out of support (nibbles follow)
        uint64_t rc[kChunk];
        int32_t ci = ci_b;
        int64_t rem = rem_b;
#pragma unroll
        for (int k = 0; k < kChunk; ++k) {
            sta[k] = 0; rg[k] = 0; rc[k] = 0;
            if (k < cnt) {
                const int32_t c = idx ? __ldg(idx + base + k) : ci;
                if (c >= 0 && c < n_cdfs) {
                    const int32_t escape = __ldg(sizes + c) - 2;
                    int32_t v = sv[k] - __ldg(offsets + c);
                    uint32_t raw = 0;
                    bool esc = false;
                    if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = escape; esc = true; }
                    else if (v >= escape) { raw = (uint32_t)(2 * (v - escape)); v = escape; esc = true; }
                    const uint32_t e = (uint32_t)c * (uint32_t)stride + (uint32_t)v;
                    const uint32_t s0 = (uint32_t)__ldg(cdfs + e), s1 = (uint32_t)__ldg(cdfs + e + 1);
                    sta[k] = s0;
                    rg[k] = (s1 > s0 ? s1 - s0 : 0u) | (esc ? 0x80000000u : 0u);
                    rc[k] = __ldg(rcp + e);
                    if (esc) sv[k] = (int32_t)raw;  // keep the raw value for phase B
                }
                if (++rem >= n_spatial) { rem = 0; ++ci; }
            }
        }
#pragma unroll
        for (int k = kChunk - 1; k >= 0; --k) {
            if (k < cnt) {
                const uint32_t range = rg[k] & 0x7fffffffu;
                if (range == 0) { bad = true; continue; }
                if (rg[k] & 0x80000000u) rans_put_escape(st, (uint32_t)sv[k]);  // rare: kept out of line
                rans_put_symbol(st, sta[k], range, rc[k]);
            }
        }
        i -= cnt;
        // the next chunk starts kChunk positions earlier
        rem_b -= kChunk;
        while (rem_b < 0) { rem_b += n_spatial; --ci_b; }
    }
    *--st.wp = st.hi;
    *--st.wp = st.lo;
    lengths[b] = bad ? -1 : (int32_t)(row + cap_words - st.wp);
}

// out[offsets[b] .. offsets[b] + lengths[b]) (words) = the used tail of scratch row b
__global__ void rans_pack_kernel(const uint32_t* __restrict__ work, int64_t cap_words, const int32_t* __restrict__ lengths,
                                 const int64_t* __restrict__ offsets, uint32_t* __restrict__ out) {
    const int b = blockIdx.x;
    const int32_t len = lengths[b];
    if (len <= 0) return;
    const uint32_t* src = work + (size_t)b * cap_words + (cap_words - len);
    uint32_t* dst = out + offsets[b];
    for (int32_t k = threadIdx.x; k < len; k += blockDim.x) dst[k] = src[k];
}

// ---------------------------------------------------------------------------------------------------------------
// decoder: one image per thread; mirrors decode_one() in host_codec.cpp
// ---------------------------------------------------------------------------------------------------------------
struct RansReader {
    const uint32_t* words;
    int64_t n_words, pos;
    uint32_t lo, hi;
    __device__ __forceinline__ uint32_t next() { return pos < n_words ? __ldg(words + pos++) : 0u; }
    __device__ __forceinline__ void refill() {  // if (x < L) x = (x << 32) | next()
        if (hi == 0 && lo < (uint32_t)kRansLow) { hi = lo; lo = next(); }
    }
    __device__ __forceinline__ uint32_t nibble() {
        const uint32_t v = lo & kRansBypassMax;
        lo = (lo >> kRansBypassBits) | (hi << (32 - kRansBypassBits));
        hi >>= kRansBypassBits;
        refill();
        return v;
    }
};

__global__ void __launch_bounds__(32) rans_decode_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ word_offsets,
                                                         const int32_t* __restrict__ n_words, const int32_t* __restrict__ indexes,
                                                         int64_t idx_stride, int batch, int64_t n, int64_t n_spatial,
                                                         const int32_t* __restrict__ cdfs, int n_cdfs, int stride,
                                                         const int32_t* __restrict__ sizes, const int32_t* __restrict__ offsets,
                                                         int32_t* __restrict__ out, int32_t* __restrict__ status) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= batch) return;
    RansReader r{packed + word_offsets[b], (int64_t)n_words[b], 0, 0u, 0u};
    if (r.n_words < 2) { status[b] = -1; return; }
    r.lo = r.next();
    r.hi = r.next();
    const int32_t* idx = indexes ? indexes + (size_t)b * idx_stride : nullptr;
    int32_t* o = out + (size_t)b * n;
    int32_t ci_run = 0;
    int64_t rem = 0;
    bool bad = false;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t ci = idx ? __ldg(idx + i) : ci_run;
        if (++rem >= n_spatial) { rem = 0; ++ci_run; }
        if (ci < 0 || ci >= n_cdfs) { bad = true; o[i] = 0; continue; }
        const int32_t* cdf = cdfs + (size_t)ci * stride;
        const int32_t len = __ldg(sizes + ci), escape = len - 2;
        const uint32_t target = r.lo & 0xffffu;
        // first entry strictly above the target (the table is strictly increasing): s = that position - 1
        int32_t lo_i = 0, hi_i = len;  // search in cdf[0 .. len)
        while (lo_i < hi_i) {
            const int32_t mid = (lo_i + hi_i) >> 1;
            if ((uint32_t)__ldg(cdf + mid) <= target) lo_i = mid + 1; else hi_i = mid;
        }
        const int32_t s = lo_i - 1;
        const uint32_t start = (uint32_t)__ldg(cdf + s), freq = (uint32_t)__ldg(cdf + s + 1) - start;
        // x = freq * (x >> 16) + (x & 0xffff) - start
        const uint64_t x = ((uint64_t)r.hi << 32) | r.lo;
        const uint64_t nx = (uint64_t)freq * (x >> kRansPrecision) + target - start;
        r.lo = (uint32_t)nx;
        r.hi = (uint32_t)(nx >> 32);
        r.refill();
        int32_t v = s;
        if (v == escape) {
            uint32_t d = r.nibble();
            uint32_t nibbles = d;
            while (d == kRansBypassMax) {
                d = r.nibble();
                nibbles += d;
            }
            uint32_t raw = 0;
            for (uint32_t j = 0; j < nibbles; ++j) {
                const uint32_t nb = r.nibble();
                if (j < 8) raw |= nb << (j * kRansBypassBits);
            }
            v = (int32_t)(raw >> 1);
            v = (raw & 1u) ? -v - 1 : v + escape;
        }
        o[i] = v + __ldg(offsets + ci);
    }
    status[b] = bad ? -1 : 0;
}


// ---------------------------------------------------------------------------------------------------------------
// Warp-per-image coders (round 2).  The one-thread-per-image kernels above spend ~640 cycles per symbol: every lane
// walks its own image, so every load is a scattered gather whose latency nothing hides (32 images per warp, 8 warps for
// a 256-tile batch).  Here ONE WARP owns an image: the 32 lanes fetch and look up a chunk of 32 symbols in parallel
// (coalesced, two chunks ahead of the coder); the coder state is replicated in every lane, a symbol's table entry reaches
// the chain by shuffles that do not depend on the state, and renormalisation is predicated (no divergent lane-0 section,
// no shared-memory hand-off: 6.0 -> 3.0 ms per 256 tiles).  Same bitstream: the arithmetic of rans_put_symbol /
// rans_put_escape / RansReader is unchanged.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRansMaxTables = 1024;  // offsets / sizes staged in shared memory

__global__ void __launch_bounds__(32) rans_encode_warp_kernel(const int32_t* __restrict__ symbols, const int32_t* __restrict__ indexes,
                                                              int64_t idx_stride, int batch, int n, int n_spatial,
                                                              const int32_t* __restrict__ cdfs, int n_cdfs, int stride,
                                                              const int32_t* __restrict__ sizes, const int32_t* __restrict__ offsets,
                                                              const uint64_t* __restrict__ rcp, uint32_t* __restrict__ work,
                                                              int64_t cap_words, int32_t* __restrict__ lengths) {
    __shared__ int32_t s_size[kRansMaxTables], s_off[kRansMaxTables];
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < n_cdfs; i += 32) { s_size[i] = sizes[i]; s_off[i] = offsets[i]; }
    __syncwarp();
    const int32_t* sym = symbols + (size_t)b * n;
    const int32_t* idx = indexes ? indexes + (size_t)b * idx_stride : nullptr;
    uint32_t* const row = work + (size_t)b * cap_words;
    RansState st{(uint32_t)kRansLow, 0u, row + cap_words};
    bool bad = false;
    const int n_chunks = (n + 31) / 32;
    // chunk t, lane j holds the (32 t + j)-th symbol in CODING order = position n - 1 - 32 t - j (the coder runs backwards)
    auto pos_of = [&](int t) { return n - 1 - 32 * t - lane; };
    // level 1: the symbol and its table index (independent loads)
    auto load_l1 = [&](int t, int32_t& sv, int32_t& c) {
        const int p = pos_of(t);
        sv = 0; c = -1;
        if (t < n_chunks && p >= 0) {
            sv = __ldg(sym + p);
            c = idx ? __ldg(idx + p) : p / n_spatial;
        }
    };
    // level 2: (start, range | escape flag, reciprocal, raw bypass value) of a level-1 pair
    auto load_l2 = [&](int32_t sv, int32_t c, uint32_t& start, uint32_t& range, uint64_t& rc, uint32_t& raw) {
        start = 0; range = 0; rc = 0; raw = 0;
        if (c >= 0 && c < n_cdfs) {
            const int32_t escape = s_size[c] - 2;
            int32_t v = sv - s_off[c];
            bool esc = false;
            if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = escape; esc = true; }
            else if (v >= escape) { raw = (uint32_t)(2 * (v - escape)); v = escape; esc = true; }
            if (escape >= 0) {
                const uint32_t e = (uint32_t)c * (uint32_t)stride + (uint32_t)v;
                const uint32_t s0 = (uint32_t)__ldg(cdfs + e), s1 = (uint32_t)__ldg(cdfs + e + 1);
                start = s0;
                range = (s1 > s0 ? s1 - s0 : 0u) | (esc ? 0x80000000u : 0u);
                rc = __ldg(rcp + e);
            }
        }
    };
    // The coder state is REPLICATED in every lane (uniform control flow, like the decoder): a symbol's table entry travels
    // from the lane that looked it up by four shuffles that do not depend on the state, so they issue ahead of the chain; no
    // shared-memory hand-off, no divergent lane-0 section, and the renormalisation is a predicated store + two selects.
    auto put = [&](uint32_t start, uint32_t range, uint32_t rc_lo, uint32_t rc_hi) {
        // Both outcomes of the renormalisation test are divided speculatively (the renormalised state is the 32-bit st.hi
        // alone), so the test itself is off the dependent chain and only selects at the end.
        const bool ren = st.hi >= (range << 15);  // x >= ((L >> 16) << 32) * range
        if (ren) *--st.wp = st.lo;                 // (every lane writes the same word to the same address)
        const uint64_t rcp = ((uint64_t)rc_hi << 32) | rc_lo;
        uint64_t qn = __umul64hi(((uint64_t)st.hi << 32) | st.lo, rcp);
        uint32_t rn = st.lo - (uint32_t)qn * range;
        if (rn >= range) { rn -= range; ++qn; }
        uint32_t qr = (uint32_t)__umul64hi((uint64_t)st.hi, rcp);  // st.hi < 2^32: the quotient fits 32 bits
        uint32_t rr = st.hi - qr * range;
        if (rr >= range) { rr -= range; ++qr; }
        const uint64_t q = ren ? (uint64_t)qr : qn;
        const uint32_t r = ren ? rr : rn;
        st.hi = (uint32_t)(q >> 16);
        st.lo = ((uint32_t)q << 16) | (r + start);
    };
    int32_t sv1, c1, sv2, c2;
    uint32_t a0, r0, w0; uint64_t q0;  // this lane's entry of the chunk being coded
    {
        int32_t sv0, c0;
        load_l1(0, sv0, c0);
        load_l1(1, sv1, c1);
        load_l2(sv0, c0, a0, r0, q0, w0);
    }
    for (int t = 0; t < n_chunks; ++t) {
        // in flight while chunk t is coded: the symbols of chunk t + 2, the table entries of chunk t + 1
        load_l1(t + 2, sv2, c2);
        uint32_t a1, r1, w1; uint64_t q1;
        load_l2(sv1, c1, a1, r1, q1, w1);
        const int cnt = min(32, n - 32 * t);
        const uint32_t rc_lo = (uint32_t)q0, rc_hi = (uint32_t)(q0 >> 32);
        const bool mine = lane < cnt;
        if (__any_sync(0xffffffffu, mine && (r0 & 0x7fffffffu) == 0u)) bad = true;  // a symbol outside its table
        if (st.wp - row < 32 * 12) bad = true;  // scratch row nearly full: the host coder takes over
        if (!bad) {
            if (cnt == 32 && !__any_sync(0xffffffffu, (r0 & 0x80000000u) != 0u)) {
#pragma unroll 8
                for (int k = 0; k < 32; ++k)
                    put(__shfl_sync(0xffffffffu, a0, k), __shfl_sync(0xffffffffu, r0, k), __shfl_sync(0xffffffffu, rc_lo, k),
                        __shfl_sync(0xffffffffu, rc_hi, k));
            } else {  // a ragged last chunk, or a chunk with out-of-support symbols (bypass coded: rare, out of line)
                for (int k = 0; k < cnt; ++k) {
                    const uint32_t rg = __shfl_sync(0xffffffffu, r0, k);
                    if (rg & 0x80000000u) rans_put_escape(st, __shfl_sync(0xffffffffu, w0, k));
                    put(__shfl_sync(0xffffffffu, a0, k), rg & 0x7fffffffu, __shfl_sync(0xffffffffu, rc_lo, k),
                        __shfl_sync(0xffffffffu, rc_hi, k));
                }
            }
        }
        a0 = a1; r0 = r1; q0 = q1; w0 = w1;
        sv1 = sv2; c1 = c2;
    }
    if (lane == 0) {
        *--st.wp = st.hi;
        *--st.wp = st.lo;
        lengths[b] = bad ? -1 : (int32_t)(row + cap_words - st.wp);
    }
}

// Decoder, index == position / n_spatial (the EntropyBottleneck layout), rows of <= 96 entries: the CDF row of the current
// channel lives in three registers per lane, the search is three compares + ballots, the word stream is prefetched 32
// words at a time and handed out by shuffles, and the coder state is replicated in every lane (uniform control flow).
__global__ void __launch_bounds__(32) rans_decode_warp_kernel(const uint32_t* __restrict__ packed, const int64_t* __restrict__ word_offsets,
                                                              const int32_t* __restrict__ n_words_all, int batch, int n, int n_spatial,
                                                              const int32_t* __restrict__ cdfs, int n_cdfs, int stride,
                                                              const int32_t* __restrict__ sizes, const int32_t* __restrict__ offsets,
                                                              int32_t* __restrict__ out, int32_t* __restrict__ status) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const uint32_t* words = packed + word_offsets[b];
    const int n_words = n_words_all[b];
    if (n_words < 2) { if (lane == 0) status[b] = -1; return; }
    // word window: lane j holds words[win0 + j]
    int win0 = 0;
    uint32_t wreg = lane < n_words ? __ldg(words + lane) : 0u;
    int pos = 0;
    auto next_word = [&]() -> uint32_t {
        if (pos - win0 >= 32) {
            win0 += 32;
            wreg = (win0 + lane) < n_words ? __ldg(words + win0 + lane) : 0u;
        }
        const uint32_t w = __shfl_sync(0xffffffffu, wreg, (pos - win0) & 31);
        const uint32_t r = pos < n_words ? w : 0u;
        ++pos;
        return r;
    };
    uint32_t lo = next_word(), hi = next_word();
    auto refill = [&]() { if (hi == 0 && lo < (uint32_t)kRansLow) { hi = lo; lo = next_word(); } };
    auto nibble = [&]() -> uint32_t {
        const uint32_t v = lo & kRansBypassMax;
        lo = (lo >> kRansBypassBits) | (hi << (32 - kRansBypassBits));
        hi >>= kRansBypassBits;
        refill();
        return v;
    };
    int32_t* o = out + (size_t)b * n;
    bool bad = false;
    int32_t outv = 0;
    uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu, r2 = 0xffffffffu;  // this lane's three entries of the current row
    int32_t escape = 0, offset = 0, len = 0;
    int i = 0;
    // One channel's symbols: straight-line per symbol (the validity checks only accumulate a flag, the refill is two selects
    // around a shuffle that is always executed); NB = registers per lane the row needs (<= 32 / 64 / 96 entries: one ballot
    // each).  (Measured and dropped: finding the interval with two compares per lane and ONE redux.sync.or of the packed
    // start / frequency instead of ballot + popc + two shuffles -- 6.2 ms against 5.35 ms per 256 tiles.)
    auto decode_run = [&](auto nb_tag, int count) {
        constexpr int NB = decltype(nb_tag)::value;
        for (int k = 0; k < count; ++k, ++i) {
            const uint32_t target = lo & 0xffffu;
            // entries <= target form a prefix of the (strictly increasing) row: their count - 1 is the symbol
            int below = __popc(__ballot_sync(0xffffffffu, r0 <= target));
            if (NB > 1) below += __popc(__ballot_sync(0xffffffffu, r1 <= target));
            if (NB > 2) below += __popc(__ballot_sync(0xffffffffu, r2 <= target));
            const int sidx = below - 1;  // below >= 1 (cdf[0] == 0); below <= len - 1 for a valid stream
            uint32_t start, next;
            if (NB == 2) {
                start = __shfl_sync(0xffffffffu, sidx < 32 ? r0 : r1, sidx & 31);
                next = __shfl_sync(0xffffffffu, below < 32 ? r0 : r1, below & 31);
            } else if (NB > 2) {
                const uint32_t cand_a = sidx < 32 ? r0 : (sidx < 64 ? r1 : r2), cand_b = below < 32 ? r0 : (below < 64 ? r1 : r2);
                start = __shfl_sync(0xffffffffu, cand_a, sidx & 31);
                next = __shfl_sync(0xffffffffu, cand_b, below & 31);
            } else {
                start = __shfl_sync(0xffffffffu, r0, sidx & 31);
                next = __shfl_sync(0xffffffffu, r0, below & 31);
            }
            const uint32_t freq = next - start;
            bad |= (below < 1) | (below >= len) | (freq == 0u) | (freq > 65536u);
            const uint64_t x = ((uint64_t)hi << 32) | lo;
            const uint64_t nx = (uint64_t)freq * (x >> kRansPrecision) + (target - start);
            const uint32_t nlo = (uint32_t)nx, nhi = (uint32_t)(nx >> 32);
            // refill: if (x < L) x = (x << 32) | next word
            if (pos - win0 >= 32) {  // (uniform, once per 32 words)
                win0 += 32;
                wreg = (win0 + lane) < n_words ? __ldg(words + win0 + lane) : 0u;
            }
            const uint32_t w = __shfl_sync(0xffffffffu, wreg, (pos - win0) & 31);
            const bool need = nhi == 0u && nlo < (uint32_t)kRansLow;
            lo = need ? (pos < n_words ? w : 0u) : nlo;
            hi = need ? nlo : nhi;
            pos += need ? 1 : 0;
            int32_t v = sidx;
            if (v == escape) {  // bypass-coded value (rare)
                uint32_t d = nibble();
                uint32_t nibbles = d;
                while (d == kRansBypassMax) {
                    d = nibble();
                    nibbles += d;
                    if (nibbles > 64) { bad = true; break; }
                }
                uint32_t raw = 0;
                for (uint32_t j = 0; j < nibbles; ++j) {
                    const uint32_t nb = nibble();
                    if (j < 8) raw |= nb << (j * kRansBypassBits);
                }
                v = (int32_t)(raw >> 1);
                v = (raw & 1u) ? -v - 1 : v + escape;
            }
            v += offset;
            if ((i & 31) == lane) outv = v;
            if ((i & 31) == 31) o[i - 31 + lane] = outv;  // one coalesced store per 32 symbols
        }
    };
    for (int c = 0; i < n; ++c) {
        const int count = min(n_spatial, n - i);
        if (c < n_cdfs && __ldg(sizes + c) >= 2) {
            len = __ldg(sizes + c);
            const int32_t* cdf = cdfs + (size_t)c * stride;
            r0 = lane < len ? (uint32_t)__ldg(cdf + lane) : 0xffffffffu;
            r1 = lane + 32 < len ? (uint32_t)__ldg(cdf + lane + 32) : 0xffffffffu;
            r2 = lane + 64 < len ? (uint32_t)__ldg(cdf + lane + 64) : 0xffffffffu;
            escape = len - 2;
            offset = __ldg(offsets + c);
            if (len <= 32) decode_run(std::integral_constant<int, 1>{}, count);
            else if (len <= 64) decode_run(std::integral_constant<int, 2>{}, count);
            else decode_run(std::integral_constant<int, 3>{}, count);
        } else {  // no table for this channel: the stream cannot be decoded; zeros out, error status
            bad = true;
            for (int k = 0; k < count; ++k, ++i) {
                if ((i & 31) == lane) outv = 0;
                if ((i & 31) == 31) o[i - 31 + lane] = outv;
            }
        }
    }
    if ((n & 31) != 0 && lane < (n & 31)) o[(n & ~31) + lane] = outv;
    if (lane == 0) status[b] = bad ? -1 : 0;
}

}  // namespace licos

using namespace licos;

extern "C" {

int licos_rans_encode_device(const int32_t* symbols, const int32_t* indexes, int64_t index_stride, int batch, int64_t n,
                             int64_t n_spatial, const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                             const int32_t* offsets, uint64_t* rcp_ws, uint32_t* work, int64_t cap_words, int32_t* lengths,
                             void* stream) {
    if (!symbols || !cdfs || !cdf_sizes || !offsets || !rcp_ws || !work || !lengths || batch < 0 || n < 0 || cap_words < 2 ||
        n_cdfs < 1 || cdf_stride < 2)
        return LICOS_ERR_INVALID;
    if (!indexes && n_spatial < 1) return LICOS_ERR_INVALID;
    if (batch == 0) return LICOS_OK;
    const int64_t entries = (int64_t)n_cdfs * cdf_stride;
    rans_rcp_kernel<<<(int)((entries + 255) / 256 < 1184 ? (entries + 255) / 256 : 1184), 256, 0, (cudaStream_t)stream>>>(
        cdfs, n_cdfs, cdf_stride, cdf_sizes, rcp_ws);
    LICOS_CUDA_OK(cudaGetLastError());
    if (n_cdfs <= kRansMaxTables && n < 0x7fffffff && n_spatial < 0x7fffffff) {
        // one warp per image: coalesced look-ups two chunks ahead of the sequential coder
        rans_encode_warp_kernel<<<batch, 32, 0, (cudaStream_t)stream>>>(symbols, indexes, index_stride, batch, (int)n,
                                                                        (int)(n_spatial > 0 ? n_spatial : 1), cdfs, n_cdfs, cdf_stride,
                                                                        cdf_sizes, offsets, rcp_ws, work, cap_words, lengths);
    } else {
        rans_encode_kernel<<<(batch + 31) / 32, 32, 0, (cudaStream_t)stream>>>(symbols, indexes, index_stride, batch, n,
                                                                                n_spatial > 0 ? n_spatial : 1, cdfs, n_cdfs,
                                                                                cdf_stride, cdf_sizes, offsets, rcp_ws, work,
                                                                                cap_words, lengths);
    }
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_rans_pack_device(const uint32_t* work, int64_t cap_words, const int32_t* lengths, const int64_t* word_offsets,
                           int batch, uint32_t* out, void* stream) {
    if (!work || !lengths || !word_offsets || !out || batch < 0) return LICOS_ERR_INVALID;
    if (batch == 0) return LICOS_OK;
    rans_pack_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(work, cap_words, lengths, word_offsets, out);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_rans_decode_device(const uint32_t* packed, const int64_t* word_offsets, const int32_t* n_words, const int32_t* indexes,
                             int64_t index_stride, int batch, int64_t n, int64_t n_spatial, const int32_t* cdfs, int n_cdfs,
                             int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* symbols, int32_t* status,
                             void* stream) {
    if (!packed || !word_offsets || !n_words || !cdfs || !cdf_sizes || !offsets || !symbols || !status || batch < 0 || n < 0 ||
        n_cdfs < 1 || cdf_stride < 2)
        return LICOS_ERR_INVALID;
    if (!indexes && n_spatial < 1) return LICOS_ERR_INVALID;
    if (batch == 0 || n == 0) return LICOS_OK;
    if (!indexes && cdf_stride <= 96 && n < 0x7fffffff && n_spatial < 0x7fffffff) {
        // the EntropyBottleneck layout: one warp per image, the channel's CDF row in registers
        rans_decode_warp_kernel<<<batch, 32, 0, (cudaStream_t)stream>>>(packed, word_offsets, n_words, batch, (int)n, (int)n_spatial,
                                                                        cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, symbols, status);
    } else {
        rans_decode_kernel<<<(batch + 31) / 32, 32, 0, (cudaStream_t)stream>>>(packed, word_offsets, n_words, indexes, index_stride,
                                                                                batch, n, n_spatial > 0 ? n_spatial : 1, cdfs,
                                                                                n_cdfs, cdf_stride, cdf_sizes, offsets, symbols, status);
    }
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

}  // extern "C"
