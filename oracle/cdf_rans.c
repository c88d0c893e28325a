/* CPU oracle, plain C: quantised CDF construction and the rANS entropy coder.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.
 *
 * Restates the published algorithms of the un-vendored dependency
 * `compressai` (>= 1.2.0, un-pinned: /root/reference/environment.yml:28-29):
 *   - compressai/cpp_exts/ops/ops.cpp          pmf_to_quantized_cdf
 *   - compressai/cpp_exts/rans/rans_interface  RansEncoder / RansDecoder
 *   - third_party/ryg_rans/rans64.h            64-bit rANS, 32-bit renorm
 * reached from the reference at eval_utils.py:201 (net.compress) and
 * eval_script.py:72,88 (net.update()).  SURVEY.md section 8a row A11 and
 * section 8f row N1 are the written spec followed here.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16
#define BYPASS_PRECISION 4
#define MAX_BYPASS_VAL ((1 << BYPASS_PRECISION) - 1)
#define RANS64_L (1ull << 31)

/* returns 0 on success, -1 on a negative / non-finite probability,
 * -2 when the rounded masses sum to zero, -3 if no frequency can be stolen. */
int oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf /* n+1 */)
{
    for (int i = 0; i < n; ++i)
        if (pmf[i] < 0 || !isfinite(pmf[i])) return -1;

    cdf[0] = 0;
    for (int i = 0; i < n; ++i)
        cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));

    int total_i = 0; /* the reference accumulates in int */
    for (int i = 0; i <= n; ++i) total_i += (int)cdf[i];
    const uint32_t total = (uint32_t)total_i;
    if (total == 0) return -2;

    for (int i = 0; i <= n; ++i)
        cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
    for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = 1u << precision;

    for (int i = 0; i < n; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            /* steal one count from the least-frequent symbol that can spare it */
            uint32_t best_freq = ~0u;
            int best_steal = -1;
            for (int j = 0; j < n; ++j) {
                uint32_t freq = cdf[j + 1] - cdf[j];
                if (freq > 1 && freq < best_freq) {
                    best_freq = freq;
                    best_steal = j;
                }
            }
            if (best_steal < 0) return -3;
            if (best_steal < i) {
                for (int j = best_steal + 1; j <= i; ++j) cdf[j]--;
            } else {
                for (int j = i + 1; j <= best_steal; ++j) cdf[j]++;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */

typedef struct {
    uint16_t start;
    uint16_t range;
    uint8_t bypass;
} sym_t;

typedef struct {
    sym_t *v;
    size_t n, cap;
} symvec_t;

static int push(symvec_t *s, uint16_t start, uint16_t range, uint8_t bypass)
{
    if (s->n == s->cap) {
        size_t cap = s->cap ? s->cap * 2 : 1024;
        sym_t *nv = (sym_t *)realloc(s->v, cap * sizeof(sym_t));
        if (!nv) return -1;
        s->v = nv;
        s->cap = cap;
    }
    s->v[s->n].start = start;
    s->v[s->n].range = range;
    s->v[s->n].bypass = bypass;
    s->n++;
    return 0;
}

static void enc_put(uint64_t *r, uint32_t **pptr, uint32_t start, uint32_t freq, uint32_t scale_bits)
{
    uint64_t x = *r;
    uint64_t x_max = ((RANS64_L >> scale_bits) << 32) * freq;
    if (x >= x_max) {
        *pptr -= 1;
        **pptr = (uint32_t)x;
        x >>= 32;
    }
    *r = ((x / freq) << scale_bits) + (x % freq) + start;
}

static void enc_put_bits(uint64_t *r, uint32_t **pptr, uint32_t val, uint32_t nbits)
{
    uint64_t x = *r;
    uint32_t freq = 1u << (16 - nbits);
    uint64_t x_max = ((RANS64_L >> 16) << 32) * freq;
    if (x >= x_max) {
        *pptr -= 1;
        **pptr = (uint32_t)x;
        x >>= 32;
    }
    *r = (x << nbits) | val;
}

/* Encodes `n` symbols.  `cdfs` is row-major [n_cdfs][cdf_stride].
 * Writes at most out_cap bytes to out; returns the byte count, or
 * -1 out of memory, -2 out buffer too small, -3 bad index. */
long oracle_rans_encode_with_indexes(const int32_t *symbols, const int32_t *indexes, long n,
                                     const int32_t *cdfs, int n_cdfs, int cdf_stride,
                                     const int32_t *cdfs_sizes, const int32_t *offsets,
                                     uint8_t *out, long out_cap)
{
    symvec_t syms = {0, 0, 0};
    for (long i = 0; i < n; ++i) {
        const int32_t cdf_idx = indexes[i];
        if (cdf_idx < 0 || cdf_idx >= n_cdfs) { free(syms.v); return -3; }
        const int32_t *cdf = cdfs + (size_t)cdf_idx * cdf_stride;
        const int32_t max_value = cdfs_sizes[cdf_idx] - 2;
        int32_t value = symbols[i] - offsets[cdf_idx];
        uint32_t raw_val = 0;
        if (value < 0) {
            raw_val = (uint32_t)(-2 * value - 1);
            value = max_value;
        } else if (value >= max_value) {
            raw_val = (uint32_t)(2 * (value - max_value));
            value = max_value;
        }
        if (push(&syms, (uint16_t)cdf[value], (uint16_t)(cdf[value + 1] - cdf[value]), 0)) { free(syms.v); return -1; }
        if (value == max_value) {
            int32_t n_bypass = 0;
            while ((raw_val >> (n_bypass * BYPASS_PRECISION)) != 0) ++n_bypass;
            int32_t val = n_bypass;
            while (val >= MAX_BYPASS_VAL) {
                if (push(&syms, MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, 1)) { free(syms.v); return -1; }
                val -= MAX_BYPASS_VAL;
            }
            if (push(&syms, (uint16_t)val, (uint16_t)(val + 1), 1)) { free(syms.v); return -1; }
            for (int32_t j = 0; j < n_bypass; ++j) {
                const int32_t v = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL;
                if (push(&syms, (uint16_t)v, (uint16_t)(v + 1), 1)) { free(syms.v); return -1; }
            }
        }
    }

    /* flush: pop in reverse, write words backwards from the end of a buffer */
    size_t words = syms.n + 2;
    uint32_t *buf = (uint32_t *)malloc(words * sizeof(uint32_t));
    if (!buf) { free(syms.v); return -1; }
    uint32_t *ptr = buf + words;
    uint64_t rans = RANS64_L;
    while (syms.n > 0) {
        sym_t s = syms.v[--syms.n];
        if (!s.bypass) enc_put(&rans, &ptr, s.start, s.range, PRECISION);
        else enc_put_bits(&rans, &ptr, s.start, BYPASS_PRECISION);
    }
    ptr -= 2;
    ptr[0] = (uint32_t)(rans >> 0);
    ptr[1] = (uint32_t)(rans >> 32);
    long nbytes = (long)((buf + words) - ptr) * (long)sizeof(uint32_t);
    long rv = nbytes;
    if (nbytes > out_cap) rv = -2;
    else memcpy(out, ptr, (size_t)nbytes);
    free(buf);
    free(syms.v);
    return rv;
}

static uint32_t dec_get_bits(uint64_t *r, const uint32_t **pptr, uint32_t n_bits)
{
    uint64_t x = *r;
    uint32_t val = (uint32_t)(x & ((1u << n_bits) - 1));
    x = x >> n_bits;
    if (x < RANS64_L) {
        x = (x << 32) | **pptr;
        *pptr += 1;
    }
    *r = x;
    return val;
}

/* Decodes n symbols from `enc` (nbytes).  Returns 0, or -3 on a bad index. */
int oracle_rans_decode_with_indexes(const uint8_t *enc, long nbytes, const int32_t *indexes, long n,
                                    const int32_t *cdfs, int n_cdfs, int cdf_stride,
                                    const int32_t *cdfs_sizes, const int32_t *offsets,
                                    int32_t *out)
{
    /* word-aligned private copy (+ slack: the decoder may read one word past a
     * short stream exactly as the reference does on its std::string) */
    size_t words = (size_t)(nbytes + 3) / 4 + 4;
    uint32_t *buf = (uint32_t *)calloc(words, sizeof(uint32_t));
    if (!buf) return -1;
    memcpy(buf, enc, (size_t)nbytes);
    const uint32_t *ptr = buf;
    uint64_t rans = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
    ptr += 2;
    const uint64_t mask = (1ull << PRECISION) - 1;

    for (long i = 0; i < n; ++i) {
        const int32_t cdf_idx = indexes[i];
        if (cdf_idx < 0 || cdf_idx >= n_cdfs) { free(buf); return -3; }
        const int32_t *cdf = cdfs + (size_t)cdf_idx * cdf_stride;
        const int32_t max_value = cdfs_sizes[cdf_idx] - 2;
        const int32_t offset = offsets[cdf_idx];
        const uint32_t cum_freq = (uint32_t)(rans & mask);
        int32_t s = 0;
        {
            const int32_t len = cdfs_sizes[cdf_idx];
            int32_t k = 0;
            while (k < len && !((uint32_t)cdf[k] > cum_freq)) ++k;
            s = k - 1;
        }
        {
            uint64_t x = rans;
            const uint32_t start = (uint32_t)cdf[s];
            const uint32_t freq = (uint32_t)(cdf[s + 1] - cdf[s]);
            x = freq * (x >> PRECISION) + (x & mask) - start;
            if (x < RANS64_L) {
                x = (x << 32) | *ptr;
                ptr += 1;
            }
            rans = x;
        }
        int32_t value = s;
        if (value == max_value) {
            int32_t val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
            int32_t n_bypass = val;
            while (val == MAX_BYPASS_VAL) {
                val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
                n_bypass += val;
            }
            int32_t raw_val = 0;
            for (int j = 0; j < n_bypass; ++j) {
                val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
                raw_val |= val << (j * BYPASS_PRECISION);
            }
            value = raw_val >> 1;
            if (raw_val & 1) value = -value - 1;
            else value += max_value;
        }
        out[i] = value + offset;
    }
    free(buf);
    return 0;
}
