// Quantisation, factorized / Gaussian likelihoods, symbol + CDF-index integer path, reductions, layout
// conversion.  All of these are HBM-bound elementwise passes: vectorised, coalesced, no tensor cores.
//
// Replaces: A8/A9 EntropyBottleneck._logits_cumulative / forward, A10
// EntropyModel.quantize / dequantize / _build_indexes, A12 GaussianConditional.forward / build_indexes,
// A13 the two reductions of RateDistortionLoss.
#include "common.cuh"

#include <string.h>

namespace licos {

// ----------------------------------------------------------------------------------------------
// EntropyBottleneck density network
// ----------------------------------------------------------------------------------------------
struct EbMeta {
    int n_layers;
    int widths[LICOS_EB_MAX_LAYERS + 1];
    int ppc;
    int form;
    float bound;
};

// logits = L_{n-1}(...L_0(x)), L_i(v) = M_i v + b_i (+ t_i * tanh(.) except for the last layer);
// M_i = softplus(_matrix_i), t_i = tanh(_factor_i) are pre-applied in `p`.
template <int W>
__device__ __forceinline__ float eb_logits(const float* __restrict__ p, const EbMeta& m, float x) {
    float cur[W], nxt[W];
#pragma unroll
    for (int k = 0; k < W; ++k) cur[k] = 0.f;
    cur[0] = x;
    int off = 0;
    for (int i = 0; i < m.n_layers; ++i) {
        const int fi = m.widths[i], fo = m.widths[i + 1];
        const float* M = p + off;
        const float* b = M + fo * fi;
        const float* t = b + fo;
        const bool last = (i == m.n_layers - 1);
#pragma unroll
        for (int o = 0; o < W; ++o) {
            float acc = 0.f;
            if (o < fo) {
#pragma unroll
                for (int k = 0; k < W; ++k)
                    if (k < fi) acc += M[o * fi + k] * cur[k];
                acc += b[o];
                if (!last) acc += t[o] * tanhf(acc);
            }
            nxt[o] = acc;
        }
#pragma unroll
        for (int k = 0; k < W; ++k) cur[k] = nxt[k];
        off += fo * fi + fo + (last ? 0 : fo);
    }
    return cur[0];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int W>
__device__ __forceinline__ float eb_likelihood(const float* __restrict__ p, const EbMeta& m, float v) {
    const float lower = eb_logits<W>(p, m, v - 0.5f);
    const float upper = eb_logits<W>(p, m, v + 0.5f);
    float lik;
    if (m.form == LICOS_EB_FORM_PLAIN) {
        lik = sigmoidf_(upper) - sigmoidf_(lower);
    } else {
        const float su = lower + upper;
        const float s = (su > 0.f) ? -1.f : ((su < 0.f) ? 1.f : 0.f);
        lik = fabsf(sigmoidf_(s * upper) - sigmoidf_(s * lower));
    }
    if (m.bound > 0.f) lik = fmaxf(lik, m.bound);
    return lik;
}

constexpr int kLutR = LICOS_EB_LUT_RADIUS;
constexpr int kLutN = 2 * kLutR + 1;

// One block per channel: likelihood of every value med + s, s in [-R, R].
template <int W>
__global__ void eb_lut_kernel(EbMeta m, const float* __restrict__ packed, const float* __restrict__ med,
                              float* __restrict__ lut) {
    extern __shared__ float sp[];
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < m.ppc; i += blockDim.x) sp[i] = packed[(size_t)c * m.ppc + i];
    __syncthreads();
    const float md = med[c];
    for (int k = threadIdx.x; k < kLutN; k += blockDim.x) {
        const float v = (float)(k - kLutR) + md;
        lut[(size_t)c * kLutN + k] = eb_likelihood<W>(sp, m, v);
    }
}

// rare path: symbol outside the table
__device__ __noinline__ float eb_likelihood_slow(const float* __restrict__ packed, EbMeta m, int c, float v) {
    return eb_likelihood<16>(packed + (size_t)c * m.ppc, m, v);
}

template <int VEC>
__global__ void eb_eval_kernel(EbMeta m, const float* __restrict__ x, const float* __restrict__ packed,
                               const float* __restrict__ med, const float* __restrict__ lut, int C, int64_t hw,
                               int64_t n_vec, float* __restrict__ y_hat, float* __restrict__ lik) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const int64_t e0 = i * VEC;
        const int c = (int)((e0 / hw) % C);
        const float md = __ldg(med + c);
        float xv[VEC], yv[VEC], lv[VEC];
        if (VEC == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(x) + i);
            xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
        } else {
            xv[0] = __ldcs(x + i);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float r = rintf(xv[j] - md);
            yv[j] = r + md;
            if (fabsf(r) <= (float)kLutR) {
                lv[j] = __ldg(lut + (size_t)c * kLutN + ((int)r + kLutR));
            } else {
                lv[j] = eb_likelihood_slow(packed, m, c, yv[j]);
            }
        }
        if (VEC == 4) {
            __stcs(reinterpret_cast<float4*>(y_hat) + i, make_float4(yv[0], yv[1], yv[2], yv[3]));
            __stcs(reinterpret_cast<float4*>(lik) + i, make_float4(lv[0], lv[1], lv[2], lv[3]));
        } else {
            __stcs(y_hat + i, yv[0]);
            __stcs(lik + i, lv[0]);
        }
    }
}

// One pass over the latent for the whole eval-mode bottleneck (A9 + A10): y_hat, likelihoods, optionally the integer
// symbols and optionally a bf16 NHWC copy of y_hat (the layout g_s / h_s read), so y is read from HBM once and the
// separate symbols and layout-conversion launches disappear.  A block owns one image and 64 consecutive positions of
// all C channels: coalesced 256-byte rows in and out for the NCHW tensors, a [C][66] bf16 tile in shared memory that
// is read back transposed (pitch 66 -> conflict-free) for 4-byte-per-thread NHWC rows.  hw % 4 == 0.
constexpr int kEbTileHw = 64, kEbTilePitch = 66;
__global__ void __launch_bounds__(256) eb_eval_tile_kernel(EbMeta m, const float* __restrict__ x,
                                                           const float* __restrict__ packed, const float* __restrict__ med,
                                                           const float* __restrict__ lut, int C, int hw,
                                                           float* __restrict__ y_hat, float* __restrict__ lik,
                                                           int32_t* __restrict__ sym, int16_t* __restrict__ sym16,
                                                           __nv_bfloat16* __restrict__ nhwc, double* __restrict__ sum_ln) {
    extern __shared__ __nv_bfloat16 tile[];  // [C][kEbTilePitch]
    __shared__ double warp_sums[8];
    double ln_acc = 0.0;
    const int tiles_per_img = (hw + kEbTileHw - 1) / kEbTileHw;
    const int b = blockIdx.x / tiles_per_img, hw0 = (blockIdx.x % tiles_per_img) * kEbTileHw;
    const int n_hw = min(kEbTileHw, hw - hw0);  // multiple of 4
    const int vecs = n_hw / 4;                   // float4 groups per channel row
    const size_t img = (size_t)b * C * hw;
    // four independent 16-byte loads in flight per thread before any dependent table gather
    constexpr int kUnroll = 4;
    for (int idx0 = threadIdx.x; idx0 < C * 16; idx0 += 256 * kUnroll) {
        float4 t[kUnroll];
        bool live[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int idx = idx0 + u * 256;
            const int c = idx >> 4, v = idx & 15;
            live[u] = idx < C * 16 && v < vecs;
            if (live[u]) t[u] = __ldcs(reinterpret_cast<const float4*>(x + img + (size_t)c * hw + hw0 + 4 * v));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (!live[u]) continue;
            const int idx = idx0 + u * 256;
            const int c = idx >> 4, v = idx & 15;
            const size_t off = img + (size_t)c * hw + hw0 + 4 * v;
            const float md = __ldg(med + c);
            const float xv[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
            float yv[4], lv[4];
            int sv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float r = rintf(xv[j] - md);
                yv[j] = r + md;
                sv[j] = (int)r;
                lv[j] = (fabsf(r) <= (float)kLutR) ? __ldg(lut + (size_t)c * kLutN + ((int)r + kLutR))
                                                   : eb_likelihood_slow(packed, m, c, yv[j]);
            }
            if (y_hat) __stcs(reinterpret_cast<float4*>(y_hat + off), make_float4(yv[0], yv[1], yv[2], yv[3]));
            if (lik) __stcs(reinterpret_cast<float4*>(lik + off), make_float4(lv[0], lv[1], lv[2], lv[3]));
            if (sym) __stcs(reinterpret_cast<int4*>(sym + off), make_int4(sv[0], sv[1], sv[2], sv[3]));
            if (sym16) {  // saturating: |symbol| > 32767 does not occur for a trained or conditioned model, and the coder's
                          // escape path would spend > 8 nibbles on it anyway
                int q[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = min(max(sv[j], -32768), 32767) & 0xffff;
                __stcs(reinterpret_cast<uint2*>(sym16 + off), make_uint2((uint32_t)q[0] | ((uint32_t)q[1] << 16),
                                                                          (uint32_t)q[2] | ((uint32_t)q[3] << 16)));
            }
            if (sum_ln) ln_acc += (double)(logf(lv[0]) + logf(lv[1])) + (double)(logf(lv[2]) + logf(lv[3]));
            if (nhwc) {
                __nv_bfloat162* row = reinterpret_cast<__nv_bfloat162*>(tile + c * kEbTilePitch + 4 * v);
                row[0] = __floats2bfloat162_rn(yv[0], yv[1]);
                row[1] = __floats2bfloat162_rn(yv[2], yv[3]);
            }
        }
    }
    if (sum_ln) {  // rate term of this block: one double atomic per block
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ln_acc += __shfl_xor_sync(0xffffffffu, ln_acc, o);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = ln_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += warp_sums[w];
            atomicAdd(sum_ln, t);
        }
    }
    if (!nhwc) return;
    __syncthreads();
    const int half_c = C / 2;  // C is even (checked by the host)
    for (int idx = threadIdx.x; idx < n_hw * half_c; idx += 256) {
        const int p = idx / half_c, cp = idx - p * half_c;
        __nv_bfloat162 v2;
        v2.x = tile[(2 * cp) * kEbTilePitch + p];
        v2.y = tile[(2 * cp + 1) * kEbTilePitch + p];
        *reinterpret_cast<__nv_bfloat162*>(nhwc + ((size_t)b * hw + hw0 + p) * C + 2 * cp) = v2;
    }
}

// Philox4x32-10, one draw per element: counter = element index, key = seed.
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t idx) {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = 0x6c69636fu, c3 = 0x73623230u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}
__device__ __forceinline__ float uniform_pm_half(uint64_t seed, uint64_t idx) {
    return (float)(philox_u32(seed, idx) >> 8) * (1.0f / 16777216.0f) - 0.5f;  // [-0.5, 0.5)
}

// blockIdx.y = channel (parameters staged in shared memory); x-blocks stride over (batch, hw).
template <int W>
__global__ void eb_noise_kernel(EbMeta m, const float* __restrict__ x, const float* __restrict__ noise,
                                uint64_t seed, const float* __restrict__ packed, int B, int C, int64_t hw,
                                float* __restrict__ y_hat, float* __restrict__ lik) {
    extern __shared__ float sp[];
    const int c = blockIdx.y;
    for (int i = threadIdx.x; i < m.ppc; i += blockDim.x) sp[i] = packed[(size_t)c * m.ppc + i];
    __syncthreads();
    const int64_t n = (int64_t)B * hw;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int64_t b = j / hw, i = j - b * hw;
        const int64_t e = (b * C + c) * hw + i;
        const float nz = noise ? __ldcs(noise + e) : uniform_pm_half(seed, (uint64_t)e);
        const float v = __ldcs(x + e) + nz;
        __stcs(y_hat + e, v);
        __stcs(lik + e, eb_likelihood<W>(sp, m, v));
    }
}

// ----------------------------------------------------------------------------------------------
// EntropyBottleneck training backward (SURVEY.md 8a rows A8, A9 differentiated; train.py:193)
// ----------------------------------------------------------------------------------------------
// One path (v - 0.5 or v + 0.5) of the density network: forward with the activations kept, then the reverse sweep.
// `acc` receives the gradient with respect to the PACKED parameters (softplus / tanh already applied), the return
// value is d logits / d input times `seed`.  W bounds the layer widths; arrays are indexed at run time (local memory):
// the latent is 1/256 of the pixels, this kernel is not bandwidth-relevant.
template <int W>
__device__ __forceinline__ float eb_logits_fwd_keep(const float* __restrict__ p, const EbMeta& m, float x,
                                                    float (&ins)[LICOS_EB_MAX_LAYERS][W], float (&ths)[LICOS_EB_MAX_LAYERS][W]) {
    float cur[W], nxt[W];
    for (int k = 0; k < W; ++k) cur[k] = 0.f;
    cur[0] = x;
    int off = 0;
    for (int i = 0; i < m.n_layers; ++i) {
        const int fi = m.widths[i], fo = m.widths[i + 1];
        const float* M = p + off;
        const float* b = M + fo * fi;
        const float* t = b + fo;
        const bool last = (i == m.n_layers - 1);
        for (int k = 0; k < fi; ++k) ins[i][k] = cur[k];
        for (int o = 0; o < fo; ++o) {
            float acc = 0.f;
            for (int k = 0; k < fi; ++k) acc += M[o * fi + k] * cur[k];
            acc += b[o];
            float th = 0.f;
            if (!last) { th = tanhf(acc); acc += t[o] * th; }
            ths[i][o] = th;
            nxt[o] = acc;
        }
        for (int k = 0; k < fo; ++k) cur[k] = nxt[k];
        off += fo * fi + fo + (last ? 0 : fo);
    }
    return cur[0];
}

template <int W>
__device__ __forceinline__ float eb_logits_bwd(const float* __restrict__ p, const EbMeta& m, float seed,
                                               const float (&ins)[LICOS_EB_MAX_LAYERS][W],
                                               const float (&ths)[LICOS_EB_MAX_LAYERS][W], float* __restrict__ acc) {
    float d_out[W], d_in[W];
    d_out[0] = seed;
    int off = m.ppc;
    for (int i = m.n_layers - 1; i >= 0; --i) {
        const int fi = m.widths[i], fo = m.widths[i + 1];
        const bool last = (i == m.n_layers - 1);
        off -= fo * fi + fo + (last ? 0 : fo);
        const float* M = p + off;
        const float* t = M + fo * fi + fo;
        float* gM = acc + off;
        float* gb = gM + fo * fi;
        float* gt = gb + fo;
        for (int k = 0; k < fi; ++k) d_in[k] = 0.f;
        for (int o = 0; o < fo; ++o) {
            float da = d_out[o];
            if (!last) {
                const float th = ths[i][o];
                gt[o] += d_out[o] * th;
                da = d_out[o] * (1.f + t[o] * (1.f - th * th));
            }
            gb[o] += da;
            for (int k = 0; k < fi; ++k) {
                gM[o * fi + k] += da * ins[i][k];
                d_in[k] += M[o * fi + k] * da;
            }
        }
        for (int k = 0; k < fi; ++k) d_out[k] = d_in[k];
    }
    return d_out[0];
}

constexpr int kEbMaxPpc = 320;  // filters (13,13,3,3) need 298

// blockIdx.y = channel.  d_x = g_yhat + g_lik * d lik / d y_hat; d_packed[c][:] += sum over the channel's elements.
template <int W>
__global__ void __launch_bounds__(128) eb_train_bwd_kernel(EbMeta m, const float* __restrict__ y_hat, const float* __restrict__ g_lik,
                                                           const float* __restrict__ g_yhat, const float* __restrict__ packed,
                                                           int B, int C, int64_t hw, float* __restrict__ d_x,
                                                           float* __restrict__ d_packed) {
    extern __shared__ float sp[];  // [ppc] parameters, [ppc] block accumulators
    float* sacc = sp + m.ppc;
    const int c = blockIdx.y;
    for (int i = threadIdx.x; i < m.ppc; i += blockDim.x) { sp[i] = packed[(size_t)c * m.ppc + i]; sacc[i] = 0.f; }
    __syncthreads();
    float acc[kEbMaxPpc];
    for (int i = 0; i < m.ppc; ++i) acc[i] = 0.f;
    float ins_u[LICOS_EB_MAX_LAYERS][W], ths_u[LICOS_EB_MAX_LAYERS][W], ins_l[LICOS_EB_MAX_LAYERS][W], ths_l[LICOS_EB_MAX_LAYERS][W];
    const int64_t n = (int64_t)B * hw;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int64_t b = j / hw, i = j - b * hw;
        const int64_t e = (b * C + c) * hw + i;
        const float v = y_hat[e];
        const float lower = eb_logits_fwd_keep<W>(sp, m, v - 0.5f, ins_l, ths_l);
        const float upper = eb_logits_fwd_keep<W>(sp, m, v + 0.5f, ins_u, ths_u);
        float g = g_lik ? g_lik[e] : 0.f;
        float su, sl, lik;  // d lik / d upper, d lik / d lower
        if (m.form == LICOS_EB_FORM_PLAIN) {
            const float a = sigmoidf_(upper), bq = sigmoidf_(lower);
            lik = a - bq;
            su = a * (1.f - a);
            sl = -bq * (1.f - bq);
        } else {
            const float sum = lower + upper;
            const float s = (sum > 0.f) ? -1.f : ((sum < 0.f) ? 1.f : 0.f);
            const float a = sigmoidf_(s * upper), bq = sigmoidf_(s * lower);
            const float d = a - bq;
            lik = fabsf(d);
            const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
            su = sg * s * a * (1.f - a);
            sl = -sg * s * bq * (1.f - bq);
        }
        if (m.bound > 0.f && !(lik >= m.bound || g < 0.f)) g = 0.f;  // LowerBound's gradient rule
        float dv = 0.f;
        if (g != 0.f) {
            dv = eb_logits_bwd<W>(sp, m, g * su, ins_u, ths_u, acc);
            dv += eb_logits_bwd<W>(sp, m, g * sl, ins_l, ths_l, acc);
        }
        d_x[e] = dv + (g_yhat ? g_yhat[e] : 0.f);
    }
    for (int i = 0; i < m.ppc; ++i) {
        float a = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[i], a);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m.ppc; i += blockDim.x) atomicAdd(d_packed + (size_t)c * m.ppc + i, sacc[i]);
}

// Fast path for the stock density network (filters (3,3,3,3): widths 1,3,3,3,3,1): every loop bound is a compile-time
// constant, so activations and the 58 parameter-gradient accumulators live in registers.
template <int NL, int F>
struct EbStatic {
    __host__ __device__ static constexpr int width(int i) { return (i == 0 || i == NL) ? 1 : F; }
    __host__ __device__ static constexpr int layer_size(int i) { return width(i + 1) * width(i) + width(i + 1) + (i == NL - 1 ? 0 : width(i + 1)); }
    __host__ __device__ static constexpr int offset(int i) { return i == 0 ? 0 : offset(i - 1) + layer_size(i - 1); }
    static constexpr int kPpc = offset(NL);
};

template <int NL, int F>
__device__ __forceinline__ float eb_static_fwd(const float* __restrict__ p, float x, float (&ins)[NL][F], float (&ths)[NL][F]) {
    float cur[F];
#pragma unroll
    for (int k = 0; k < F; ++k) cur[k] = 0.f;
    cur[0] = x;
    int off = 0;  // a running literal after unrolling
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int fi = (i == 0) ? 1 : F, fo = (i == NL - 1) ? 1 : F;
        float nxt[F];
#pragma unroll
        for (int k = 0; k < F; ++k) ins[i][k] = cur[k];
#pragma unroll
        for (int o = 0; o < F; ++o) {
            float acc = 0.f, th = 0.f;
            if (o < fo) {
#pragma unroll
                for (int k = 0; k < F; ++k)
                    if (k < fi) acc += p[off + o * fi + k] * cur[k];
                acc += p[off + fo * fi + o];
                if (i < NL - 1) { th = tanhf(acc); acc += p[off + fo * fi + fo + o] * th; }
            }
            ths[i][o] = th;
            nxt[o] = acc;
        }
#pragma unroll
        for (int k = 0; k < F; ++k) cur[k] = nxt[k];
        off += fo * fi + fo + (i < NL - 1 ? fo : 0);
    }
    return cur[0];
}

template <int NL, int F>
__device__ __forceinline__ float eb_static_bwd(const float* __restrict__ p, float seed, const float (&ins)[NL][F],
                                               const float (&ths)[NL][F], float (&acc)[EbStatic<NL, F>::kPpc]) {
    using S = EbStatic<NL, F>;
    float d_out[F];
#pragma unroll
    for (int k = 0; k < F; ++k) d_out[k] = 0.f;
    d_out[0] = seed;
    int off = S::kPpc;
#pragma unroll
    for (int i = NL - 1; i >= 0; --i) {
        const int fi = (i == 0) ? 1 : F, fo = (i == NL - 1) ? 1 : F;
        off -= fo * fi + fo + (i < NL - 1 ? fo : 0);
        float d_in[F];
#pragma unroll
        for (int k = 0; k < F; ++k) d_in[k] = 0.f;
#pragma unroll
        for (int o = 0; o < F; ++o) {
            if (o < fo) {
                float da = d_out[o];
                if (i < NL - 1) {
                    const float th = ths[i][o];
                    acc[off + fo * fi + fo + o] += d_out[o] * th;
                    da = d_out[o] * (1.f + p[off + fo * fi + fo + o] * (1.f - th * th));
                }
                acc[off + fo * fi + o] += da;
#pragma unroll
                for (int k = 0; k < F; ++k)
                    if (k < fi) {
                        acc[off + o * fi + k] += da * ins[i][k];
                        d_in[k] += p[off + o * fi + k] * da;
                    }
            }
        }
#pragma unroll
        for (int k = 0; k < F; ++k) d_out[k] = d_in[k];
    }
    return d_out[0];
}

// Training-mode forward for the stock density network: the same arithmetic, in the same order, as eb_logits<3>, with
// compile-time layer shapes (no per-layer offset arithmetic or width predicates).
template <int NL, int F>
__device__ __forceinline__ float eb_static_logits(const float* __restrict__ p, float x) {
    float cur[F];
#pragma unroll
    for (int k = 0; k < F; ++k) cur[k] = 0.f;
    cur[0] = x;
    int off = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int fi = (i == 0) ? 1 : F, fo = (i == NL - 1) ? 1 : F;
        float nxt[F];
#pragma unroll
        for (int o = 0; o < F; ++o) {
            float acc = 0.f;
            if (o < fo) {
#pragma unroll
                for (int k = 0; k < F; ++k)
                    if (k < fi) acc += p[off + o * fi + k] * cur[k];
                acc += p[off + fo * fi + o];
                if (i < NL - 1) acc += p[off + fo * fi + fo + o] * tanhf(acc);
            }
            nxt[o] = acc;
        }
#pragma unroll
        for (int k = 0; k < F; ++k) cur[k] = nxt[k];
        off += fo * fi + fo + (i < NL - 1 ? fo : 0);
    }
    return cur[0];
}

template <int NL, int F>
__global__ void __launch_bounds__(256) eb_noise_static_kernel(EbMeta m, const float* __restrict__ x, const float* __restrict__ noise,
                                                              uint64_t seed, const float* __restrict__ packed, int B, int C, int64_t hw,
                                                              float* __restrict__ y_hat, float* __restrict__ lik) {
    constexpr int kPpc = EbStatic<NL, F>::kPpc;
    __shared__ float sp[kPpc];
    const int c = blockIdx.y;
    for (int i = threadIdx.x; i < kPpc; i += blockDim.x) sp[i] = packed[(size_t)c * kPpc + i];
    __syncthreads();
    const int64_t n = (int64_t)B * hw;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int64_t b = j / hw, i = j - b * hw;
        const int64_t e = (b * C + c) * hw + i;
        const float nz = noise ? __ldcs(noise + e) : uniform_pm_half(seed, (uint64_t)e);
        const float v = __ldcs(x + e) + nz;
        const float lower = eb_static_logits<NL, F>(sp, v - 0.5f), upper = eb_static_logits<NL, F>(sp, v + 0.5f);
        float l;
        if (m.form == LICOS_EB_FORM_PLAIN) {
            l = sigmoidf_(upper) - sigmoidf_(lower);
        } else {
            const float su = lower + upper;
            const float s = (su > 0.f) ? -1.f : ((su < 0.f) ? 1.f : 0.f);
            l = fabsf(sigmoidf_(s * upper) - sigmoidf_(s * lower));
        }
        if (m.bound > 0.f) l = fmaxf(l, m.bound);
        __stcs(y_hat + e, v);
        __stcs(lik + e, l);
    }
}

template <int NL, int F>
__global__ void __launch_bounds__(128) eb_train_bwd_static_kernel(EbMeta m, const float* __restrict__ y_hat,
                                                                  const float* __restrict__ g_lik, const float* __restrict__ g_yhat,
                                                                  const float* __restrict__ packed, int B, int C, int64_t hw,
                                                                  float* __restrict__ d_x, float* __restrict__ d_packed) {
    constexpr int kPpc = EbStatic<NL, F>::kPpc;
    __shared__ float sp[kPpc], sacc[kPpc];
    const int c = blockIdx.y;
    for (int i = threadIdx.x; i < kPpc; i += blockDim.x) { sp[i] = packed[(size_t)c * kPpc + i]; sacc[i] = 0.f; }
    __syncthreads();
    float acc[kPpc];
#pragma unroll
    for (int i = 0; i < kPpc; ++i) acc[i] = 0.f;
    const int64_t n = (int64_t)B * hw;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int64_t b = j / hw, i = j - b * hw;
        const int64_t e = (b * C + c) * hw + i;
        const float v = y_hat[e];
        float ins_u[NL][F], ths_u[NL][F], ins_l[NL][F], ths_l[NL][F];
        const float lower = eb_static_fwd<NL, F>(sp, v - 0.5f, ins_l, ths_l);
        const float upper = eb_static_fwd<NL, F>(sp, v + 0.5f, ins_u, ths_u);
        float g = g_lik ? g_lik[e] : 0.f;
        float su, sl, lik;
        if (m.form == LICOS_EB_FORM_PLAIN) {
            const float a = sigmoidf_(upper), bq = sigmoidf_(lower);
            lik = a - bq;
            su = a * (1.f - a);
            sl = -bq * (1.f - bq);
        } else {
            const float sum = lower + upper;
            const float s = (sum > 0.f) ? -1.f : ((sum < 0.f) ? 1.f : 0.f);
            const float a = sigmoidf_(s * upper), bq = sigmoidf_(s * lower);
            const float d = a - bq;
            lik = fabsf(d);
            const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
            su = sg * s * a * (1.f - a);
            sl = -sg * s * bq * (1.f - bq);
        }
        if (m.bound > 0.f && !(lik >= m.bound || g < 0.f)) g = 0.f;
        float dv = eb_static_bwd<NL, F>(sp, g * su, ins_u, ths_u, acc);
        dv += eb_static_bwd<NL, F>(sp, g * sl, ins_l, ths_l, acc);
        d_x[e] = dv + (g_yhat ? g_yhat[e] : 0.f);
    }
#pragma unroll
    for (int i = 0; i < kPpc; ++i) {
        float a = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[i], a);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPpc; i += blockDim.x) atomicAdd(d_packed + (size_t)c * kPpc + i, sacc[i]);
}

// ----------------------------------------------------------------------------------------------
// Parameter block <-> raw parameters of the density network, and the auxiliary (quantile) loss
// ----------------------------------------------------------------------------------------------
struct EbRaw {
    const float* matrix[LICOS_EB_MAX_LAYERS];
    const float* bias[LICOS_EB_MAX_LAYERS];
    const float* factor[LICOS_EB_MAX_LAYERS];
};
struct EbRawOut {
    float* matrix[LICOS_EB_MAX_LAYERS];
    float* bias[LICOS_EB_MAX_LAYERS];
    float* factor[LICOS_EB_MAX_LAYERS];
};

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch's threshold 20

// MODE 0: packed[c][:] = (softplus(_matrix_i), _bias_i, tanh(_factor_i))_i
// MODE 1: raw gradients from d_packed: d_matrix = d * sigmoid(_matrix), d_bias = d, d_factor = d * (1 - tanh(_factor)^2)
template <int MODE>
__global__ void eb_pack_kernel(EbMeta m, EbRaw raw, EbRawOut out, int C, float* __restrict__ packed) {
    const int total = C * m.ppc;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int c = e / m.ppc;
        int r = e - c * m.ppc;
        for (int i = 0; i < m.n_layers; ++i) {
            const int fi = m.widths[i], fo = m.widths[i + 1];
            const bool last = (i == m.n_layers - 1);
            if (r < fo * fi) {
                const size_t j = (size_t)c * fo * fi + r;
                if (MODE == 0) packed[e] = softplus_t(raw.matrix[i][j]);
                else out.matrix[i][j] = packed[e] * sigmoidf_(raw.matrix[i][j]);
                break;
            }
            r -= fo * fi;
            if (r < fo) {
                const size_t j = (size_t)c * fo + r;
                if (MODE == 0) packed[e] = raw.bias[i][j];
                else out.bias[i][j] = packed[e];
                break;
            }
            r -= fo;
            if (!last) {
                if (r < fo) {
                    const size_t j = (size_t)c * fo + r;
                    const float th = tanhf(raw.factor[i][j]);
                    if (MODE == 0) packed[e] = th;
                    else out.factor[i][j] = packed[e] * (1.f - th * th);
                    break;
                }
                r -= fo;
            }
        }
    }
}

// EntropyBottleneck.loss(): sum |logits_cumulative(quantiles, stop_gradient=True) - target|; one thread per (channel, k).
// loss[0] += the sum, d_quantiles[c][k] = sign(logit - target_k) * d logit / d quantile.
template <int W>
__global__ void eb_aux_loss_kernel(EbMeta m, const float* __restrict__ packed, const float* __restrict__ quantiles,
                                   const float* __restrict__ target, int C, float* __restrict__ loss, float* __restrict__ d_quantiles) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (e < C * 3) {
        const int c = e / 3, k = e - 3 * c;
        const float* p = packed + (size_t)c * m.ppc;
        float ins[LICOS_EB_MAX_LAYERS][W], ths[LICOS_EB_MAX_LAYERS][W];
        const float logit = eb_logits_fwd_keep<W>(p, m, quantiles[e], ins, ths);
        const float diff = logit - target[k];
        part = fabsf(diff);
        // d logit / d input: reverse sweep through the layers without the parameter gradients
        float d_out[W], d_in[W];
        d_out[0] = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        int off = m.ppc;
        for (int i = m.n_layers - 1; i >= 0; --i) {
            const int fi = m.widths[i], fo = m.widths[i + 1];
            const bool last = (i == m.n_layers - 1);
            off -= fo * fi + fo + (last ? 0 : fo);
            const float* M = p + off;
            const float* t = M + fo * fi + fo;
            for (int q = 0; q < fi; ++q) d_in[q] = 0.f;
            for (int o = 0; o < fo; ++o) {
                const float da = last ? d_out[o] : d_out[o] * (1.f + t[o] * (1.f - ths[i][o] * ths[i][o]));
                for (int q = 0; q < fi; ++q) d_in[q] += M[o * fi + q] * da;
            }
            for (int q = 0; q < fi; ++q) d_out[q] = d_in[q];
        }
        d_quantiles[e] = d_out[0];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(loss, part);
}

__global__ void eb_symbols_kernel(const float* __restrict__ x, const float* __restrict__ med, int C, int64_t hw,
                                  int64_t n, int32_t* __restrict__ sym, int32_t* __restrict__ idx) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = (int)((i / hw) % C);
        sym[i] = (int32_t)rintf(__ldcs(x + i) - __ldg(med + c));
        if (idx) idx[i] = c;
    }
}

__global__ void eb_dequant_kernel(const int32_t* __restrict__ sym, const float* __restrict__ med, int C,
                                  int64_t hw, int64_t n, float* __restrict__ y_hat) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = (int)((i / hw) % C);
        y_hat[i] = (float)sym[i] + __ldg(med + c);
    }
}

// ----------------------------------------------------------------------------------------------
// GaussianConditional
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float std_cumulative(float v) { return 0.5f * erfcf(-0.70710678118654752440f * v); }

__global__ void gc_forward_kernel(const float* __restrict__ y, const float* __restrict__ scales,
                                  const float* __restrict__ means, const float* __restrict__ noise, uint64_t seed,
                                  int64_t n, int training, float scale_bound, float lik_bound,
                                  float* __restrict__ y_hat, float* __restrict__ lik) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float mu = means ? __ldcs(means + i) : 0.f;
        const float yi = __ldcs(y + i);
        float out;
        if (training) {
            const float nz = noise ? __ldcs(noise + i) : uniform_pm_half(seed, (uint64_t)i);
            out = yi + nz;
        } else if (means) {
            out = rintf(yi - mu) + mu;
        } else {
            out = rintf(yi);
        }
        const float val = fabsf(means ? out - mu : out);
        const float s = fmaxf(__ldcs(scales + i), scale_bound);
        const float upper = std_cumulative((0.5f - val) / s);
        const float lower = std_cumulative((-0.5f - val) / s);
        float l = upper - lower;
        if (lik_bound > 0.f) l = fmaxf(l, lik_bound);
        __stcs(y_hat + i, out);
        __stcs(lik + i, l);
    }
}

// Backward of GaussianConditional.forward(y, scales, means, training=True) (SURVEY.md 8a row A12 differentiated):
//   v = |y_hat - mu|, s = max(scale, bound), lik = Phi((.5 - v)/s) - Phi((-.5 - v)/s)
//   d lik/d v = (phi(b) - phi(a)) / s,  d lik/d s = (b phi(b) - a phi(a)) / s,  a = (.5 - v)/s, b = (-.5 - v)/s
// with LowerBound's gradient rule on the likelihood (1e-9) and on the scale (0.11).  16 B in, 8-12 B out per element.
__global__ void gc_backward_kernel(const float* __restrict__ y_hat, const float* __restrict__ scales, const float* __restrict__ means,
                                   const float* __restrict__ g_lik, const float* __restrict__ g_yhat, int64_t n, float scale_bound,
                                   float lik_bound, float* __restrict__ d_y, float* __restrict__ d_scales, float* __restrict__ d_means) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float mu = means ? __ldcs(means + i) : 0.f;
        const float diff = __ldcs(y_hat + i) - mu;
        const float val = fabsf(diff);
        const float sc = __ldcs(scales + i);
        const float s = fmaxf(sc, scale_bound);
        const float a = (0.5f - val) / s, b = (-0.5f - val) / s;
        const float lik = std_cumulative(a) - std_cumulative(b);
        float g = g_lik ? __ldcs(g_lik + i) : 0.f;
        if (lik_bound > 0.f && !(lik >= lik_bound || g < 0.f)) g = 0.f;
        const float pa = 0.3989422804014327f * expf(-0.5f * a * a), pb = 0.3989422804014327f * expf(-0.5f * b * b);
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float dv = g * (pb - pa) / s * sgn;
        float ds = g * (b * pb - a * pa) / s;
        if (!(sc >= scale_bound || ds < 0.f)) ds = 0.f;
        __stcs(d_y + i, dv + (g_yhat ? __ldcs(g_yhat + i) : 0.f));
        __stcs(d_scales + i, ds);
        if (d_means) __stcs(d_means + i, -dv);
    }
}

__global__ void gc_indexes_kernel(const float* __restrict__ scales, int64_t n, const float* __restrict__ table,
                                  int n_table, float scale_bound, int32_t* __restrict__ idx) {
    extern __shared__ float st[];
    for (int k = threadIdx.x; k < n_table; k += blockDim.x) st[k] = table[k];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float s = fmaxf(__ldcs(scales + i), scale_bound);
        int v = n_table - 1;
        for (int k = 0; k < n_table - 1; ++k) v -= (s <= st[k]) ? 1 : 0;
        idx[i] = v;
    }
}

__global__ void gc_symbols_kernel(const float* __restrict__ y, const float* __restrict__ means, int64_t n,
                                  int32_t* __restrict__ sym) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = means ? __ldcs(y + i) - __ldcs(means + i) : __ldcs(y + i);
        sym[i] = (int32_t)rintf(v);
    }
}

// ----------------------------------------------------------------------------------------------
// Reductions
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_accumulate(double v, double* acc) {
    __shared__ double warp_sums[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_sums[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(acc, v);
    }
}

__global__ void sum_log_kernel(const float* __restrict__ lik, int64_t n, double* acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double s = 0.0;
    float part = 0.f;
    int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        part += logf(__ldcs(lik + i));
        if (++cnt == 32) { s += (double)part; part = 0.f; cnt = 0; }
    }
    s += (double)part;
    block_accumulate(s, acc);
}

__global__ void sum_sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double s = 0.0;
    float part = 0.f;
    int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = __ldcs(a + i) - __ldcs(b + i);
        part += d * d;
        if (++cnt == 32) { s += (double)part; part = 0.f; cnt = 0; }
    }
    s += (double)part;
    block_accumulate(s, acc);
}

// backward of the two reductions of RateDistortionLoss: out = coef * g[0] / lik  and  out = coef * g[0] * (a - b);
// g is a DEVICE scalar (the incoming gradient of the bpp / mse term), so nothing synchronises.
__global__ void scaled_reciprocal_kernel(const float* __restrict__ lik, int64_t n, float coef, const float* __restrict__ g,
                                         float* __restrict__ out) {
    const float k = coef * g[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = k / lik[i];
}
__global__ void scaled_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float coef,
                                   const float* __restrict__ g, float* __restrict__ out) {
    const float k = coef * g[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = k * (a[i] - b[i]);
}

__global__ void weighted_sum2_kernel(const float* __restrict__ a, const float* __restrict__ b, float wa, float wb,
                                     int64_t n, float* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // two roundings, like `wa * a` followed by `+= wb * b` in the reference
        const float t = __fmul_rn(wa, a[i]);
        dst[i] = __fadd_rn(t, __fmul_rn(wb, b[i]));
    }
}
__global__ void scale_kernel(float* __restrict__ buf, float w, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] *= w;
}

// ----------------------------------------------------------------------------------------------
// Layout conversion: per image a [C][HW] <-> [HW][C] transpose through a 32x33 shared tile
// ----------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C,
                                         int64_t hw, int take_abs) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float* src = in + (size_t)b * C * hw;
    __nv_bfloat16* dst = out + (size_t)b * C * hw;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r;
        const int64_t p = p0 + threadIdx.x;
        float v = 0.f;
        if (c < C && p < hw) v = src[(size_t)c * hw + p];
        tile[r][threadIdx.x] = take_abs ? fabsf(v) : v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t p = p0 + r;
        const int c = c0 + threadIdx.x;
        if (c < C && p < hw) dst[(size_t)p * C + c] = __float2bfloat16_rn(tile[threadIdx.x][r]);
    }
}

__global__ void nhwc_bf16_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int C,
                                         int64_t hw) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const __nv_bfloat16* src = in + (size_t)b * C * hw;
    float* dst = out + (size_t)b * C * hw;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t p = p0 + r;
        const int c = c0 + threadIdx.x;
        float v = 0.f;
        if (c < C && p < hw) v = __bfloat162float(src[(size_t)p * C + c]);
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r;
        const int64_t p = p0 + threadIdx.x;
        if (c < C && p < hw) dst[(size_t)c * hw + p] = tile[threadIdx.x][r];
    }
}

static int grid_for(int64_t n, int threads, int per_sm = 8) {
    int64_t g = (n + threads - 1) / threads;
    const int64_t cap = 148LL * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

static bool make_meta(const licos_eb_params* p, EbMeta& m, int& max_w) {
    if (!p || p->n_layers < 1 || p->n_layers > LICOS_EB_MAX_LAYERS || p->channels < 1 || !p->packed || !p->medians)
        return false;
    m.n_layers = p->n_layers;
    max_w = 1;
    int ppc = 0;
    for (int i = 0; i <= p->n_layers; ++i) {
        m.widths[i] = p->widths[i];
        if (p->widths[i] < 1 || p->widths[i] > 16) return false;
        if (p->widths[i] > max_w) max_w = p->widths[i];
    }
    for (int i = p->n_layers + 1; i <= LICOS_EB_MAX_LAYERS; ++i) m.widths[i] = 0;
    if (p->widths[0] != 1 || p->widths[p->n_layers] != 1) return false;
    for (int i = 0; i < p->n_layers; ++i) {
        ppc += p->widths[i + 1] * p->widths[i] + p->widths[i + 1];
        if (i < p->n_layers - 1) ppc += p->widths[i + 1];
    }
    if (ppc != p->params_per_channel) return false;
    m.ppc = ppc;
    m.form = p->form;
    m.bound = p->likelihood_bound;
    return true;
}

}  // namespace licos

using namespace licos;

extern "C" {

int64_t licos_eb_lut_floats(int channels) { return (int64_t)channels * kLutN; }

int licos_eb_forward_eval(const licos_eb_params* p, const float* x, int batch, int64_t hw, float* lut_ws,
                          float* y_hat, float* lik, void* stream) {
    EbMeta m;
    int max_w;
    if (!make_meta(p, m, max_w) || !x || !lut_ws || !y_hat || !lik || batch < 0 || hw < 0) return LICOS_ERR_INVALID;
    const int64_t n = (int64_t)batch * p->channels * hw;
    if (n == 0) return LICOS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t sm = (size_t)m.ppc * sizeof(float);
    if (max_w <= 3) eb_lut_kernel<3><<<p->channels, 288, sm, s>>>(m, p->packed, p->medians, lut_ws);
    else eb_lut_kernel<16><<<p->channels, 288, sm, s>>>(m, p->packed, p->medians, lut_ws);
    LICOS_CUDA_OK(cudaGetLastError());
    const bool vec = (hw % 4 == 0) && (((uintptr_t)x | (uintptr_t)y_hat | (uintptr_t)lik) % 16 == 0);
    if (vec) {
        const int64_t nv = n / 4;
        eb_eval_kernel<4><<<grid_for(nv, 256), 256, 0, s>>>(m, x, p->packed, p->medians, lut_ws, p->channels, hw, nv,
                                                           y_hat, lik);
    } else {
        eb_eval_kernel<1><<<grid_for(n, 256), 256, 0, s>>>(m, x, p->packed, p->medians, lut_ws, p->channels, hw, n,
                                                          y_hat, lik);
    }
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_build_lut(const licos_eb_params* p, float* lut, void* stream) {
    EbMeta m;
    int max_w;
    if (!make_meta(p, m, max_w) || !lut) return LICOS_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t sm = (size_t)m.ppc * sizeof(float);
    if (max_w <= 3) eb_lut_kernel<3><<<p->channels, 288, sm, s>>>(m, p->packed, p->medians, lut);
    else eb_lut_kernel<16><<<p->channels, 288, sm, s>>>(m, p->packed, p->medians, lut);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_eval_fused(const licos_eb_params* p, const licos_eb_fused_args* a, void* stream) {
    EbMeta m;
    int max_w;
    if (!a || !make_meta(p, m, max_w) || !a->x || !a->lut || a->batch < 0 || a->hw < 0) return LICOS_ERR_INVALID;
    if (!a->y_hat && !a->lik && !a->symbols && !a->symbols_i16 && !a->y_hat_nhwc_bf16 && !a->sum_ln) return LICOS_ERR_INVALID;
    const int64_t hw = a->hw;
    if (hw % 4 != 0 || hw > 0x7fffffff || p->channels % 2 != 0 || p->channels > 512) return LICOS_ERR_UNSUPPORTED;
    if ((((uintptr_t)a->x | (uintptr_t)a->y_hat | (uintptr_t)a->lik | (uintptr_t)a->symbols) % 16) != 0 ||
        ((uintptr_t)a->symbols_i16 % 8) != 0)
        return LICOS_ERR_UNSUPPORTED;
    const int64_t n = (int64_t)a->batch * p->channels * hw;
    if (n == 0) return LICOS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (!a->lut_ready) {
        const int rc = licos_eb_build_lut(p, a->lut, stream);
        if (rc != LICOS_OK) return rc;
    }
    const int64_t blocks = (int64_t)a->batch * ((hw + kEbTileHw - 1) / kEbTileHw);
    if (blocks > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    const size_t tile_bytes = a->y_hat_nhwc_bf16 ? (size_t)p->channels * kEbTilePitch * sizeof(__nv_bfloat16) : 0;
    LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)eb_eval_tile_kernel, 96 * 1024));
    eb_eval_tile_kernel<<<(int)blocks, 256, tile_bytes, s>>>(m, a->x, p->packed, p->medians, a->lut, p->channels, (int)hw,
                                                            a->y_hat, a->lik, a->symbols, a->symbols_i16,
                                                            (__nv_bfloat16*)a->y_hat_nhwc_bf16, a->sum_ln);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_forward_noise(const licos_eb_params* p, const float* x, const float* noise, uint64_t seed, int batch,
                           int64_t hw, float* y_hat, float* lik, void* stream) {
    EbMeta m;
    int max_w;
    if (!make_meta(p, m, max_w) || !x || !y_hat || !lik || batch < 0 || hw < 0) return LICOS_ERR_INVALID;
    const int64_t n = (int64_t)batch * hw;
    if (n == 0) return LICOS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t sm = (size_t)m.ppc * sizeof(float);
    int gx = (int)((n + 255) / 256);
    const int cap = (148 * 8 + p->channels - 1) / p->channels;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, p->channels);
    bool stock = m.n_layers == 5 && m.ppc == EbStatic<5, 3>::kPpc;
    for (int i = 1; i < 5 && stock; ++i) stock = m.widths[i] == 3;
    if (stock) eb_noise_static_kernel<5, 3><<<grid, 256, 0, s>>>(m, x, noise, seed, p->packed, batch, p->channels, hw, y_hat, lik);
    else if (max_w <= 3) eb_noise_kernel<3><<<grid, 256, sm, s>>>(m, x, noise, seed, p->packed, batch, p->channels, hw, y_hat, lik);
    else eb_noise_kernel<16><<<grid, 256, sm, s>>>(m, x, noise, seed, p->packed, batch, p->channels, hw, y_hat, lik);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_backward(const licos_eb_params* p, const float* y_hat, const float* g_lik, const float* g_yhat, int batch,
                      int64_t hw, float* d_x, float* d_packed, void* stream) {
    EbMeta m;
    int max_w;
    if (!make_meta(p, m, max_w) || !y_hat || !d_x || !d_packed || batch < 0 || hw < 0) return LICOS_ERR_INVALID;
    if (m.ppc > kEbMaxPpc) return LICOS_ERR_UNSUPPORTED;
    const int64_t n = (int64_t)batch * hw;
    if (n == 0) return LICOS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t sm = (size_t)m.ppc * 2 * sizeof(float);
    int gx = (int)((n + 127) / 128);
    const int cap = (148 * 8 + p->channels - 1) / p->channels;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, p->channels);
    bool stock = m.n_layers == 5 && m.ppc == EbStatic<5, 3>::kPpc;
    for (int i = 1; i < 5 && stock; ++i) stock = m.widths[i] == 3;
    if (stock) eb_train_bwd_static_kernel<5, 3><<<grid, 128, 0, s>>>(m, y_hat, g_lik, g_yhat, p->packed, batch, p->channels, hw, d_x, d_packed);
    else if (max_w <= 3) eb_train_bwd_kernel<3><<<grid, 128, sm, s>>>(m, y_hat, g_lik, g_yhat, p->packed, batch, p->channels, hw, d_x, d_packed);
    else eb_train_bwd_kernel<16><<<grid, 128, sm, s>>>(m, y_hat, g_lik, g_yhat, p->packed, batch, p->channels, hw, d_x, d_packed);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

static bool make_meta_shape(int channels, int n_layers, const int* widths, EbMeta& m, int& max_w) {
    if (channels < 1 || n_layers < 1 || n_layers > LICOS_EB_MAX_LAYERS || !widths) return false;
    m.n_layers = n_layers;
    max_w = 1;
    int ppc = 0;
    for (int i = 0; i <= n_layers; ++i) {
        m.widths[i] = widths[i];
        if (widths[i] < 1 || widths[i] > 16) return false;
        if (widths[i] > max_w) max_w = widths[i];
    }
    for (int i = n_layers + 1; i <= LICOS_EB_MAX_LAYERS; ++i) m.widths[i] = 0;
    if (widths[0] != 1 || widths[n_layers] != 1) return false;
    for (int i = 0; i < n_layers; ++i) ppc += widths[i + 1] * widths[i] + widths[i + 1] + (i < n_layers - 1 ? widths[i + 1] : 0);
    m.ppc = ppc;
    m.form = 0;
    m.bound = 0.f;
    return true;
}

int licos_eb_pack_params(const licos_eb_raw_params* raw, int channels, int n_layers, const int* widths, float* packed,
                         void* stream) {
    EbMeta m;
    int max_w;
    if (!raw || !packed || !make_meta_shape(channels, n_layers, widths, m, max_w)) return LICOS_ERR_INVALID;
    EbRaw r;
    EbRawOut o;
    memset(&o, 0, sizeof(o));
    for (int i = 0; i < LICOS_EB_MAX_LAYERS; ++i) { r.matrix[i] = raw->matrix[i]; r.bias[i] = raw->bias[i]; r.factor[i] = raw->factor[i]; }
    for (int i = 0; i < n_layers; ++i)
        if (!r.matrix[i] || !r.bias[i] || (i < n_layers - 1 && !r.factor[i])) return LICOS_ERR_INVALID;
    eb_pack_kernel<0><<<grid_for((int64_t)channels * m.ppc, 256), 256, 0, (cudaStream_t)stream>>>(m, r, o, channels, packed);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_param_grads(const licos_eb_raw_params* raw, const float* d_packed, int channels, int n_layers, const int* widths,
                         const licos_eb_raw_grads* grads, void* stream) {
    EbMeta m;
    int max_w;
    if (!raw || !d_packed || !grads || !make_meta_shape(channels, n_layers, widths, m, max_w)) return LICOS_ERR_INVALID;
    EbRaw r;
    EbRawOut o;
    for (int i = 0; i < LICOS_EB_MAX_LAYERS; ++i) {
        r.matrix[i] = raw->matrix[i]; r.bias[i] = raw->bias[i]; r.factor[i] = raw->factor[i];
        o.matrix[i] = grads->matrix[i]; o.bias[i] = grads->bias[i]; o.factor[i] = grads->factor[i];
    }
    for (int i = 0; i < n_layers; ++i)
        if (!r.matrix[i] || !r.bias[i] || !o.matrix[i] || !o.bias[i] || (i < n_layers - 1 && (!r.factor[i] || !o.factor[i])))
            return LICOS_ERR_INVALID;
    eb_pack_kernel<1><<<grid_for((int64_t)channels * m.ppc, 256), 256, 0, (cudaStream_t)stream>>>(m, r, o, channels,
                                                                                                const_cast<float*>(d_packed));
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_aux_loss(const float* packed, int channels, int n_layers, const int* widths, const float* quantiles,
                      const float* target, float* loss, float* d_quantiles, void* stream) {
    EbMeta m;
    int max_w;
    if (!packed || !quantiles || !target || !loss || !d_quantiles || !make_meta_shape(channels, n_layers, widths, m, max_w))
        return LICOS_ERR_INVALID;
    const int n = channels * 3;
    if (max_w <= 3) eb_aux_loss_kernel<3><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(m, packed, quantiles, target, channels, loss, d_quantiles);
    else eb_aux_loss_kernel<16><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(m, packed, quantiles, target, channels, loss, d_quantiles);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_symbols(const float* x, const float* medians, int batch, int channels, int64_t hw, int32_t* symbols,
                     int32_t* indexes, void* stream) {
    if (!x || !medians || !symbols || batch < 0 || channels < 1 || hw < 0) return LICOS_ERR_INVALID;
    const int64_t n = (int64_t)batch * channels * hw;
    if (n == 0) return LICOS_OK;
    eb_symbols_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, medians, channels, hw, n, symbols, indexes);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_eb_dequantize(const int32_t* symbols, const float* medians, int batch, int channels, int64_t hw,
                        float* y_hat, void* stream) {
    if (!symbols || !medians || !y_hat || batch < 0 || channels < 1 || hw < 0) return LICOS_ERR_INVALID;
    const int64_t n = (int64_t)batch * channels * hw;
    if (n == 0) return LICOS_OK;
    eb_dequant_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(symbols, medians, channels, hw, n, y_hat);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gc_forward(const float* y, const float* scales, const float* means, const float* noise, uint64_t seed,
                     int64_t n, int training, float scale_bound, float likelihood_bound, float* y_hat, float* lik,
                     void* stream) {
    if (!y || !scales || !y_hat || !lik || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    gc_forward_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(y, scales, means, noise, seed, n, training,
                                                                        scale_bound, likelihood_bound, y_hat, lik);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gc_backward(const float* y_hat, const float* scales, const float* means, const float* g_lik, const float* g_yhat,
                      int64_t n, float scale_bound, float likelihood_bound, float* d_y, float* d_scales, float* d_means,
                      void* stream) {
    if (!y_hat || !scales || !d_y || !d_scales || n < 0 || (d_means && !means)) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    gc_backward_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(y_hat, scales, means, g_lik, g_yhat, n, scale_bound,
                                                                          likelihood_bound, d_y, d_scales, d_means);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gc_build_indexes(const float* scales, int64_t n, const float* table, int n_table, float scale_bound,
                           int32_t* indexes, void* stream) {
    if (!scales || !table || !indexes || n < 0 || n_table < 1 || n_table > 4096) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    gc_indexes_kernel<<<grid_for(n, 256), 256, n_table * sizeof(float), (cudaStream_t)stream>>>(
        scales, n, table, n_table, scale_bound, indexes);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gc_symbols(const float* y, const float* means, int64_t n, int32_t* symbols, void* stream) {
    if (!y || !symbols || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    gc_symbols_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(y, means, n, symbols);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_sum_log(const float* lik, int64_t n, double* acc, void* stream) {
    if (!lik || !acc || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    sum_log_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(lik, n, acc);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_sum_sq_err(const float* a, const float* b, int64_t n, double* acc, void* stream) {
    if (!a || !b || !acc || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    sum_sq_err_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(a, b, n, acc);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_scaled_reciprocal(const float* lik, int64_t n, float coef, const float* g_dev, float* out, void* stream) {
    if (!lik || !g_dev || !out || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    scaled_reciprocal_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(lik, n, coef, g_dev, out);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_scaled_diff(const float* a, const float* b, int64_t n, float coef, const float* g_dev, float* out, void* stream) {
    if (!a || !b || !g_dev || !out || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    scaled_diff_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, coef, g_dev, out);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_weighted_sum2(const float* a, const float* b, float w_a, float w_b, int64_t n, float* dst, void* stream) {
    if (!a || !b || !dst || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    weighted_sum2_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, w_a, w_b, n, dst);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_scale_inplace(float* buf, float w, int64_t n, void* stream) {
    if (!buf || n < 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    scale_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(buf, w, n);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_nchw_f32_to_nhwc_bf16(const float* in, void* out, int batch, int channels, int64_t hw, int take_abs,
                                void* stream) {
    if (!in || !out || batch < 0 || channels < 1 || hw < 0) return LICOS_ERR_INVALID;
    if (batch == 0 || hw == 0) return LICOS_OK;
    if (batch > 65535) return LICOS_ERR_UNSUPPORTED;
    dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((channels + 31) / 32), (unsigned)batch);
    nchw_to_nhwc_bf16_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, channels, hw,
                                                                           take_abs);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_nhwc_bf16_to_nchw_f32(const void* in, float* out, int batch, int channels, int64_t hw, void* stream) {
    if (!in || !out || batch < 0 || channels < 1 || hw < 0) return LICOS_ERR_INVALID;
    if (batch == 0 || hw == 0) return LICOS_OK;
    if (batch > 65535) return LICOS_ERR_UNSUPPORTED;
    dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((channels + 31) / 32), (unsigned)batch);
    nhwc_bf16_to_nchw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, out, channels,
                                                                           hw);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

}  // extern "C"
