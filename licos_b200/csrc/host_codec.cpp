// Host-side integer path: quantised CDF tables and the rANS coder (bitstream-compatible with
// compressai.ans), plus the small ABI utilities.  Pure C++17, no CUDA.
//
// Replaces compressai._CXX.pmf_to_quantized_cdf and compressai.ans.RansEncoder/RansDecoder
// (reached from eval_script.py:72,88 `net.update()` and eval_utils.py:201 `net.compress`).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/licos_b200.h"

namespace {

constexpr uint32_t kPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr uint32_t kBypassMax = (1u << kBypassBits) - 1;
constexpr uint64_t kRansL = 1ull << 31;

thread_local int g_last_cuda_error = 0;

struct Tables {
    const int32_t* cdfs;
    int n_cdfs;
    int stride;
    const int32_t* sizes;
    const int32_t* offsets;
};

// One coding step: either a modelled symbol (start, range out of 2^16) or a 4-bit raw nibble.
struct Step {
    uint16_t start;
    uint16_t range;  // 0 marks a raw nibble stored in `start`
};

class WordStack {
   public:
    explicit WordStack(size_t cap) : buf_(cap), pos_(cap) {}
    void push(uint32_t w) { buf_[--pos_] = w; }
    const uint32_t* data() const { return buf_.data() + pos_; }
    size_t words() const { return buf_.size() - pos_; }

   private:
    std::vector<uint32_t> buf_;
    size_t pos_;
};

int64_t encode_one(const int32_t* symbols, const int32_t* indexes, int64_t n, const Tables& t, uint8_t* out,
                   int64_t cap) {
    std::vector<Step> steps;
    steps.reserve((size_t)n + 16);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        if (ci < 0 || ci >= t.n_cdfs) return LICOS_ERR_INVALID;
        const int32_t* cdf = t.cdfs + (size_t)ci * t.stride;
        const int32_t escape = t.sizes[ci] - 2;  // last slot of the table = "out of range"
        int32_t v = symbols[i] - t.offsets[ci];
        uint32_t raw = 0;
        if (v < 0) {
            raw = (uint32_t)(-2 * v - 1);
            v = escape;
        } else if (v >= escape) {
            raw = (uint32_t)(2 * (v - escape));
            v = escape;
        }
        steps.push_back({(uint16_t)cdf[v], (uint16_t)(cdf[v + 1] - cdf[v])});
        if (v == escape) {
            uint32_t nibbles = 0;
            while (nibbles < 8 && (raw >> (nibbles * kBypassBits)) != 0) ++nibbles;  // (a shift by 32 is undefined)
            uint32_t count = nibbles;  // nibble count, base-15 "unary" prefix
            while (count >= kBypassMax) {
                steps.push_back({(uint16_t)kBypassMax, 0});
                count -= kBypassMax;
            }
            steps.push_back({(uint16_t)count, 0});
            for (uint32_t j = 0; j < nibbles; ++j)
                steps.push_back({(uint16_t)((raw >> (j * kBypassBits)) & kBypassMax), 0});
        }
    }

    WordStack ws(steps.size() + 2);
    uint64_t x = kRansL;
    for (size_t k = steps.size(); k-- > 0;) {
        const Step s = steps[k];
        if (s.range != 0) {
            const uint64_t x_max = ((kRansL >> kPrecision) << 32) * s.range;
            if (x >= x_max) {
                ws.push((uint32_t)x);
                x >>= 32;
            }
            x = ((x / s.range) << kPrecision) + (x % s.range) + s.start;
        } else {
            const uint64_t x_max = ((kRansL >> 16) << 32) * (uint64_t)(1u << (16 - kBypassBits));
            if (x >= x_max) {
                ws.push((uint32_t)x);
                x >>= 32;
            }
            x = (x << kBypassBits) | s.start;
        }
    }
    ws.push((uint32_t)(x >> 32));
    ws.push((uint32_t)x);
    const int64_t nbytes = (int64_t)ws.words() * 4;
    if (nbytes > cap) return LICOS_ERR_BUFFER;
    std::memcpy(out, ws.data(), (size_t)nbytes);
    return nbytes;
}

struct Reader {
    std::vector<uint32_t> words;
    size_t pos = 0;
    uint64_t x = 0;
    uint32_t next() { return pos < words.size() ? words[pos++] : 0u; }
    uint32_t nibble() {
        const uint32_t v = (uint32_t)(x & kBypassMax);
        x >>= kBypassBits;
        if (x < kRansL) x = (x << 32) | next();
        return v;
    }
};

int decode_one(const uint8_t* enc, int64_t nbytes, const int32_t* indexes, int64_t n, const Tables& t,
               int32_t* out) {
    if (nbytes < 8) return LICOS_ERR_INVALID;
    Reader r;
    r.words.resize((size_t)(nbytes + 3) / 4, 0u);
    std::memcpy(r.words.data(), enc, (size_t)nbytes);
    r.x = (uint64_t)r.next();
    r.x |= (uint64_t)r.next() << 32;
    const uint64_t mask = (1ull << kPrecision) - 1;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        if (ci < 0 || ci >= t.n_cdfs) return LICOS_ERR_INVALID;
        const int32_t* cdf = t.cdfs + (size_t)ci * t.stride;
        const int32_t len = t.sizes[ci];
        const int32_t escape = len - 2;
        const uint32_t target = (uint32_t)(r.x & mask);
        // first entry strictly above the target (the table is strictly increasing)
        const int32_t* it = std::upper_bound(cdf, cdf + len, (int32_t)target);
        const int32_t s = (int32_t)(it - cdf) - 1;
        const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
        r.x = (uint64_t)freq * (r.x >> kPrecision) + (r.x & mask) - start;
        if (r.x < kRansL) r.x = (r.x << 32) | r.next();
        int32_t v = s;
        if (v == escape) {
            uint32_t d = r.nibble();
            uint32_t nibbles = d;
            while (d == kBypassMax) {
                d = r.nibble();
                nibbles += d;
            }
            uint32_t raw = 0;
            for (uint32_t j = 0; j < nibbles; ++j) raw |= r.nibble() << (j * kBypassBits);
            v = (int32_t)(raw >> 1);
            v = (raw & 1u) ? -v - 1 : v + escape;
        }
        out[i] = v + t.offsets[ci];
    }
    return LICOS_OK;
}

template <class F>
int parallel_for(int count, int threads, F&& body) {
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > count) threads = count;
    std::atomic<int> next{0};
    std::atomic<int> status{LICOS_OK};
    auto worker = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= count) break;
            const int rc = body(i);
            if (rc != LICOS_OK) status.store(rc);
        }
    };
    if (threads <= 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        pool.reserve((size_t)threads);
        for (int k = 0; k < threads; ++k) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    return status.load();
}

}  // namespace

extern "C" {

void licos_set_last_cuda_error(int code) { g_last_cuda_error = code; }
int licos_last_cuda_error(void) { return g_last_cuda_error; }
int licos_abi_version(void) { return LICOS_ABI_VERSION; }

const char* licos_strerror(int code) {
    switch (code) {
        case LICOS_OK: return "ok";
        case LICOS_ERR_INVALID: return "invalid argument";
        case LICOS_ERR_CUDA: return "CUDA call failed (see licos_last_cuda_error)";
        case LICOS_ERR_UNSUPPORTED: return "unsupported configuration";
        case LICOS_ERR_NO_DEVICE: return "no sm_100 (B200) device";
        case LICOS_ERR_DOMAIN: return "pmf has a negative, non-finite or all-zero mass";
        case LICOS_ERR_NOMEM: return "host allocation failed";
        case LICOS_ERR_BUFFER: return "output buffer too small";
        default: return "unknown error";
    }
}

int licos_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf) {
    if (!pmf || !cdf || n < 1 || precision < 1 || precision > 31) return LICOS_ERR_INVALID;
    for (int i = 0; i < n; ++i)
        if (!(pmf[i] >= 0.f) || !std::isfinite(pmf[i])) return LICOS_ERR_DOMAIN;

    const float scale = (float)(1 << precision);
    cdf[0] = 0;
    int sum = 0;  // the reference sums in `int`
    for (int i = 0; i < n; ++i) {
        cdf[i + 1] = (uint32_t)std::round(pmf[i] * scale);
        sum += (int)cdf[i + 1];
    }
    const uint32_t total = (uint32_t)sum;
    if (total == 0) return LICOS_ERR_DOMAIN;

    // rescale each mass to the target precision, then prefix-sum
    uint32_t run = 0;
    for (int i = 0; i <= n; ++i) {
        run += (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
        cdf[i] = run;
    }
    cdf[n] = 1u << precision;

    // every symbol needs a non-zero slot: borrow from the cheapest donor that can spare one
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;
        uint32_t donor_freq = std::numeric_limits<uint32_t>::max();
        int donor = -1;
        for (int j = 0; j < n; ++j) {
            const uint32_t f = cdf[j + 1] - cdf[j];
            if (f > 1 && f < donor_freq) {
                donor_freq = f;
                donor = j;
            }
        }
        if (donor < 0) return LICOS_ERR_DOMAIN;
        if (donor < i) {
            for (int j = donor + 1; j <= i; ++j) --cdf[j];
        } else {
            for (int j = i + 1; j <= donor; ++j) ++cdf[j];
        }
    }
    return LICOS_OK;
}

int64_t licos_rans_encode(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs, int n_cdfs,
                          int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out,
                          int64_t out_capacity) {
    if (n < 0 || (n > 0 && (!symbols || !indexes)) || !cdfs || !cdf_sizes || !offsets || !out) return LICOS_ERR_INVALID;
    try {
        return encode_one(symbols, indexes, n, Tables{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets}, out, out_capacity);
    } catch (const std::bad_alloc&) {
        return LICOS_ERR_NOMEM;
    }
}

int licos_rans_decode(const uint8_t* encoded, int64_t n_bytes, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                      int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* symbols) {
    if (!encoded || n < 0 || (n > 0 && (!indexes || !symbols)) || !cdfs || !cdf_sizes || !offsets) return LICOS_ERR_INVALID;
    try {
        return decode_one(encoded, n_bytes, indexes, n, Tables{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets}, symbols);
    } catch (const std::bad_alloc&) {
        return LICOS_ERR_NOMEM;
    }
}

int licos_rans_encode_batch(const int32_t* symbols, const int32_t* indexes, int batch, int64_t n,
                            int64_t index_batch_stride, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out, int64_t out_stride,
                            int64_t* out_sizes, int threads) {
    if (batch < 0 || n < 0 || !symbols || !indexes || !cdfs || !cdf_sizes || !offsets || !out || !out_sizes)
        return LICOS_ERR_INVALID;
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    try {
        return parallel_for(batch, threads, [&](int i) -> int {
            const int64_t rc = encode_one(symbols + (size_t)i * n, indexes + (size_t)i * index_batch_stride, n, t,
                                          out + (size_t)i * out_stride, out_stride);
            if (rc < 0) return (int)rc;
            out_sizes[i] = rc;
            return LICOS_OK;
        });
    } catch (const std::bad_alloc&) {
        return LICOS_ERR_NOMEM;
    }
}

int licos_rans_decode_batch(const uint8_t* const* encoded, const int64_t* n_bytes, const int32_t* indexes, int batch,
                            int64_t n, int64_t index_batch_stride, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, int32_t* symbols, int threads) {
    if (batch < 0 || n < 0 || !encoded || !n_bytes || !indexes || !cdfs || !cdf_sizes || !offsets || !symbols)
        return LICOS_ERR_INVALID;
    const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    try {
        return parallel_for(batch, threads, [&](int i) -> int {
            return decode_one(encoded[i], n_bytes[i], indexes + (size_t)i * index_batch_stride, n, t,
                              symbols + (size_t)i * n);
        });
    } catch (const std::bad_alloc&) {
        return LICOS_ERR_NOMEM;
    }
}

}  // extern "C"
