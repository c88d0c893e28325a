// Backward-pass kernels of the training step (SURVEY.md section 8a rows A3-A6 differentiated; reached from
// licos/train.py:193 `out_criterion["loss"].backward()`).
//
//   wgrad_kernel        weight gradient of conv 5x5 s2 / deconv 5x5 s2 / conv 3x3 s1 / 1x1 (and the GDN gamma gradient,
//                       which is a 1x1 weight gradient): out[tap][cs][cb] += sum over pixels of
//                       small[pixel][cs] * big[pixel shifted by the tap][cb].  The contraction runs over PIXELS, so both
//                       operands are MN-major UMMA operands read straight from the NHWC bf16 tiles TMA delivers
//                       (channels contiguous = the M / N dimension).  A CTA keeps up to four taps (four 128 x 128 fp32
//                       accumulators = all of TMEM) that share one parity view of `big`: one TMA slab with a one-pixel
//                       halo serves all of them as shifted operand views.  Work is split over (tap group, channel
//                       block, pixel range) units; a unit ends with a vector red.add flush into the fp32 result.
//   gdn_bwd_* / relu_bwd / colsum_bf16   the elementwise and reduction pieces around the 1x1 channel mixes of the GDN
//                       backward pass (the mixes themselves run on the conv engine as LICOS_CONV_1X1 layers).
//
// The data gradients (dgrad) need no kernel of their own: dgrad of a stride-2 conv IS the transposed conv with the
// same weight and vice versa, so they run on conv_engine.cu with re-packed weights.
#include "common.cuh"
#include "gdn_bwd.cuh"

#include <stdlib.h>
#include <type_traits>
#include <string.h>
#include <mutex>

namespace licos {

constexpr int kWgTH = 8, kWgTW = 16;            // tile of `small`: 8 rows x 16 columns = 128 pixels = 8 K-steps of 16
constexpr int kWgSlabRows = 9, kWgSlabCols = 18;  // slab of `big`: the tile + one row and two columns of halo
constexpr uint32_t kWgSChunk = 128u * 128u;     // [128 pixels][64 channels] bf16
constexpr uint32_t kWgGBytes = (uint32_t)kWgSlabRows * kWgSlabCols * 128u;  // 20 736 B delivered per slab chunk
constexpr uint32_t kWgGChunk = 21u * 1024u;     // slab chunk pitch (1 KB aligned)
constexpr uint32_t kWgStage = 2u * kWgSChunk + 2u * kWgGChunk;
constexpr int kWgStages = 3;
constexpr int kWgMaxTaps = 4, kWgMaxGroups = 10, kWgMaxCombos = 96;
constexpr int kWgThreads = 192;  // producer, MMA issuer, 4 flush warps

struct WgTap {
    int8_t dr, dc;   // slab row / column of the tile origin for this tap
    int16_t w_tap;   // index of the [cs][cb] matrix in `out`
};
struct WgGroup {
    int8_t view, r0, n_taps, pad_;
    WgTap taps[kWgMaxTaps];
};

struct WgradParams {
    CUtensorMap s_map;
    CUtensorMap g_maps[4];
    WgGroup groups[kWgMaxGroups];
    int n_groups, m_blocks, n_blocks;
    int Cs, Cb;
    int tiles_h, tiles_w, n_tiles;
    int n_combos, n_units;
    int unit_begin[kWgMaxCombos + 1];
    float* out;
};

struct WgUnit {
    int g, mb, nb, t0, t1;
};
__device__ __forceinline__ WgUnit wg_decode(const WgradParams& p, int u) {
    int c = 0;
    while (c + 1 < p.n_combos && u >= p.unit_begin[c + 1]) ++c;
    const int s = u - p.unit_begin[c], ns = p.unit_begin[c + 1] - p.unit_begin[c];
    WgUnit w;
    w.nb = c % p.n_blocks;
    w.mb = (c / p.n_blocks) % p.m_blocks;
    w.g = c / (p.n_blocks * p.m_blocks);
    w.t0 = (int)((long long)s * p.n_tiles / ns);
    w.t1 = (int)((long long)(s + 1) * p.n_tiles / ns);
    return w;
}

// MN-major SWIZZLE_128B operand: 64 channels (128 B) contiguous per pixel row, 8-pixel groups 1 KB apart,
// the next 64 channels `lbo_bytes` further on.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[kWgStages], empty[kWgStages], acc_full, acc_empty;
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full, 1);
        mbar_init(&acc_empty, 128);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int tiles_per_img = p.tiles_h * p.tiles_w;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== producer =====================
            tma_prefetch_desc(&p.s_map);
            for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.g_maps[i]);
            uint32_t slot = 0, phase = 0;
            for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                const WgUnit w = wg_decode(p, u);
                const WgGroup& g = p.groups[w.g];
                const int n_g = (p.Cb - w.nb * 128) > 64 ? 2 : 1;
                const uint32_t bytes = 2u * kWgSChunk + (uint32_t)n_g * kWgGBytes;
                for (int t = w.t0; t < w.t1; ++t) {
                    const int b = t / tiles_per_img, r = t % tiles_per_img;
                    const int h0 = (r / p.tiles_w) * kWgTH, w0 = (r % p.tiles_w) * kWgTW;
                    mbar_wait(&empty[slot], phase ^ 1u);
                    uint8_t* st = smem + (size_t)slot * kWgStage;
                    mbar_arrive_expect_tx(&full[slot], bytes);
                    for (int j = 0; j < 2; ++j)
                        tma_load_4d(st + j * kWgSChunk, &p.s_map, &full[slot], (w.mb * 2 + j) * 64, w0, h0, b);
                    for (int j = 0; j < n_g; ++j)
                        tma_load_4d(st + 2 * kWgSChunk + j * kWgGChunk, &p.g_maps[g.view], &full[slot], (w.nb * 2 + j) * 64,
                                    w0 - 1, h0 + g.r0, b);
                    if (++slot == kWgStages) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
        const uint64_t a_hi = umma_desc_mn_sw128(kWgSChunk), b_hi = umma_desc_mn_sw128(kWgGChunk);
        uint32_t slot = 0, phase = 0, n_unit = 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++n_unit) {
            const WgUnit w = wg_decode(p, u);
            const WgGroup& g = p.groups[w.g];
            const int nt = (p.Cb - w.nb * 128) > 64 ? 128 : 64;
            const uint32_t idesc = umma_idesc_bf16(128, nt) | (1u << 15) | (1u << 16);
            const int n_taps = g.n_taps;
            uint32_t b_off[kWgMaxTaps];
#pragma unroll
            for (int k = 0; k < kWgMaxTaps; ++k)
                b_off[k] = (uint32_t)(((int)g.taps[k].dr * kWgSlabCols + (int)g.taps[k].dc) * 128) >> 4;
            mbar_wait_warp(&acc_empty, (n_unit & 1u) ^ 1u);
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int t = w.t0; t < w.t1; ++t) {
                mbar_wait_warp(&full[slot], phase);
                tc_fence_after();
                const uint32_t st16 = (smem_base + slot * kWgStage) >> 4;
                const uint32_t g16 = st16 + ((2u * kWgSChunk) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (uint32_t r = 0; r < (uint32_t)kWgTH; ++r) {
                        const uint64_t ad = a_hi | (uint64_t)(st16 + r * ((kWgTW * 128u) >> 4));
                        const uint32_t brow = g16 + r * ((kWgSlabCols * 128u) >> 4);
#pragma unroll
                        for (int k = 0; k < kWgMaxTaps; ++k)
                            if (k < n_taps)
                                umma_bf16(tmem_base + (uint32_t)k * 128u, ad, b_hi | (uint64_t)(brow + b_off[k]), idesc,
                                          accumulate | r);
                    }
                    umma_commit(&empty[slot]);
                }
                __syncwarp();
                accumulate = 1;
                if (++slot == kWgStages) { slot = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(&acc_full);
            __syncwarp();
        }
    } else {
        // ===================== flush: TMEM -> red.add into the fp32 result =====================
        const uint32_t q = (uint32_t)(warp & 3);
        const uint32_t lane_sel = (q * 32u) << 16;
        const int row = (int)q * 32 + lane;  // TMEM lane == channel of `small` inside the block
        uint32_t n_unit = 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++n_unit) {
            const WgUnit w = wg_decode(p, u);
            const WgGroup& g = p.groups[w.g];
            const int nt = (p.Cb - w.nb * 128) > 64 ? 128 : 64;
            const int n_valid = p.Cb - w.nb * 128;  // columns of this block that exist in `out` (a multiple of 16)
            const int cs = w.mb * 128 + row;
            mbar_wait(&acc_full, n_unit & 1u);
            tc_fence_after();
            for (int k = 0; k < g.n_taps; ++k) {
                float* o = p.out + ((size_t)g.taps[k].w_tap * p.Cs + cs) * p.Cb + w.nb * 128;
                for (int cc = 0; cc < nt / 32; ++cc) {
                    float v[32];
                    tmem_ld32(tmem_base + lane_sel + (uint32_t)k * 128u + cc * 32, v);
                    tmem_ld_wait();
                    if (cs < p.Cs) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (cc * 32 + 4 * j < n_valid) red_add_v4(o + cc * 32 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace licos
#include "wgrad_image.cuh"
namespace licos {

// ----------------------------------------------------------------------------------------------
// elementwise / reduction kernels of the GDN, ReLU and bias backward passes (bf16 NHWC, 8 channels per thread)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

__global__ void square_bf16_kernel(const uint4* __restrict__ x, uint4* __restrict__ x2, int64_t n8) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float a[8];
        unpack8(__ldg(x + i), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] *= a[j];
        x2[i] = pack8(a);
    }
}

// Per-channel sums of what an elementwise kernel just produced (beta / bias gradients) without a second pass: the
// launch makes gridDim.x * blockDim.x a multiple of C / 8, so a thread always sees the same 8 channels.
__device__ __forceinline__ void colsum_flush(const float (&s)[8], int cg, int C, float* part, float* __restrict__ acc) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) part[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&part[cg * 8 + j], s[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(acc + i, part[i]);
}

// norm = beta + gamma . x^2 (from the 1x1 layer).  GDN: y = x * rsqrt(norm); IGDN: y = x * sqrt(norm).
//   d_direct = g * dy/dx at fixed norm;  d_norm = g * dy/dnorm;  sum_dn[c] += column sums of d_norm (= beta_hat.grad)
template <bool INVERSE>
__global__ void __launch_bounds__(256) gdn_bwd_mid_kernel(const uint4* __restrict__ x, const uint4* __restrict__ g,
                                                          const uint4* __restrict__ norm, uint4* __restrict__ d_norm,
                                                          uint4* __restrict__ d_direct, int64_t n8, int C, float* __restrict__ sum_dn) {
    __shared__ float part[512];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = i0; i < n8; i += stride) {
        float a[8], b[8], c[8], dn[8], dd[8];
        unpack8(__ldg(x + i), a);
        unpack8(__ldg(g + i), b);
        unpack8(__ldg(norm + i), c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (INVERSE) {
                const float sq = sqrtf(c[j]);
                dd[j] = b[j] * sq;
                dn[j] = 0.5f * b[j] * a[j] / sq;
            } else {
                const float r = rsqrtf(c[j]);
                dd[j] = b[j] * r;
                dn[j] = -0.5f * b[j] * a[j] * r * r * r;
            }
        }
        const uint4 pk = pack8(dn);
        d_norm[i] = pk;
        d_direct[i] = pack8(dd);
        if (sum_dn) {  // sum what the weight-gradient kernel will read: the bf16-rounded values
            float r8[8];
            unpack8(pk, r8);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += r8[j];
        }
    }
    if (sum_dn) colsum_flush(s, (int)(i0 % (C / 8)), C, part, sum_dn);
}

// dx = d_direct + 2 x t, t = gamma^T . d_norm (from the 1x1 layer); in place over d_direct is allowed;
// sum_dx[c] += column sums of dx (= the conv's bias.grad)
__global__ void __launch_bounds__(256) gdn_bwd_out_kernel(const uint4* __restrict__ x, const uint4* __restrict__ t, const uint4* d_direct,
                                                          uint4* dx, int64_t n8, int C, float* __restrict__ sum_dx) {
    __shared__ float part[512];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = i0; i < n8; i += stride) {
        float a[8], b[8], c[8];
        unpack8(__ldg(x + i), a);
        unpack8(__ldg(t + i), b);
        unpack8(d_direct[i], c);
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = fmaf(2.f * a[j], b[j], c[j]);
        const uint4 pk = pack8(c);
        dx[i] = pk;
        if (sum_dx) {
            float r8[8];
            unpack8(pk, r8);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += r8[j];
        }
    }
    if (sum_dx) colsum_flush(s, (int)(i0 % (C / 8)), C, part, sum_dx);
}

// Gradient through NonNegativeParametrizer (p_hat = max(p, bound)^2 - pedestal) with LowerBound's rule (A6):
// d_lb = d_hat * 2 max(p, bound); it passes where p >= bound or where it would move p back above the bound (d_lb < 0).
__global__ void gdn_param_grad_kernel(const float* __restrict__ beta, const float* __restrict__ gamma,
                                      const float* __restrict__ d_beta_hat, const float* __restrict__ d_gamma_hat, int C, float bb,
                                      float gb, float* __restrict__ d_beta, float* __restrict__ d_gamma) {
    const int64_t total = (int64_t)C * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const float gm = gamma[e];
        const float dl = d_gamma_hat[e] * 2.f * fmaxf(gm, gb);
        d_gamma[e] = (gm >= gb || dl < 0.f) ? dl : 0.f;
        if (e < C) {
            const float bt = beta[e];
            const float db = d_beta_hat[e] * 2.f * fmaxf(bt, bb);
            d_beta[e] = (bt >= bb || db < 0.f) ? db : 0.f;
        }
    }
}

__global__ void relu_bwd_kernel(const uint4* __restrict__ y, const uint4* g, uint4* dx, int64_t n8) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float a[8], b[8];
        unpack8(__ldg(y + i), a);
        unpack8(g[i], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = a[j] > 0.f ? b[j] : 0.f;
        dx[i] = pack8(b);
    }
}

// acc[c] += sum over rows of x[row][c]; x bf16 [rows][C], C % 8 == 0, C <= 512.  One thread = 8 channels of a row.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const uint4* __restrict__ x, int64_t rows, int C, float* __restrict__ acc) {
    __shared__ float part[512];
    for (int i = threadIdx.x; i < C; i += blockDim.x) part[i] = 0.f;
    __syncthreads();
    const int g_per_row = C / 8;
    const int rows_per_it = blockDim.x / g_per_row;
    const int my_g = threadIdx.x % g_per_row, my_r = threadIdx.x / g_per_row;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (my_r < rows_per_it) {
        for (int64_t r = (int64_t)blockIdx.x * rows_per_it + my_r; r < rows; r += (int64_t)gridDim.x * rows_per_it) {
            float a[8];
            unpack8(__ldg(x + r * g_per_row + my_g), a);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += a[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&part[my_g * 8 + j], s[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(acc + i, part[i]);
}

// rows[(b, oh, ow)][k] = x[b][c][2*oh + kh - 2][2*ow + kw - 2], k = (c*5 + kh)*5 + kw, zero padded to k_pad: the patch
// matrix of a 5x5 stride-2 window over a fp32 NCHW tensor (first-layer / last-layer weight gradients).  A block owns
// 64 consecutive output pixels of one row: the C x 5 x 131 input patch is read once, coalesced, into shared memory;
// every thread then assembles 16-byte groups of 8 k's through a k -> patch-offset table and writes them coalesced.
constexpr int kImPix = 64, kImPitch = 2 * kImPix + 4;
__global__ void __launch_bounds__(256) im2col5x5s2_kernel(const float* __restrict__ x, int B, int C, int H, int W, int OH, int OW,
                                                          int k_pad, int segs, __nv_bfloat16* __restrict__ rows) {
    extern __shared__ float im_smem[];
    float* patch = im_smem;                                              // [C][5][kImPitch]
    int16_t* tab = reinterpret_cast<int16_t*>(patch + C * 5 * kImPitch);  // [k_pad]
    const int K = C * 25;
    int blk = blockIdx.x;
    const int seg = blk % segs;
    blk /= segs;
    const int oh = blk % OH, b = blk / OH;
    const int ow0 = seg * kImPix;
    for (int k = threadIdx.x; k < k_pad; k += blockDim.x)
        tab[k] = (int16_t)(k < K ? ((k / 25) * 5 + (k % 25) / 5) * kImPitch + (k % 5) : -1);
    for (int idx = threadIdx.x; idx < C * 5 * kImPitch; idx += blockDim.x) {
        const int col = idx % kImPitch, r = (idx / kImPitch) % 5, c = idx / (5 * kImPitch);
        const int ih = 2 * oh + r - 2, iw = 2 * ow0 + col - 2;
        patch[idx] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(x + (((size_t)b * C + c) * H + ih) * W + iw) : 0.f;
    }
    __syncthreads();
    const int groups = k_pad / 8;
    for (int item = threadIdx.x; item < kImPix * groups; item += blockDim.x) {
        const int pl = item / groups, gk = item % groups;
        if (ow0 + pl >= OW) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int t = tab[gk * 8 + j];
            v[j] = t < 0 ? 0.f : patch[t + 2 * pl];
        }
        *reinterpret_cast<uint4*>(rows + (((size_t)b * OH + oh) * OW + ow0 + pl) * k_pad + gk * 8) = pack8(v);
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    });
    return fn;
}
static bool wg_make_map(CUtensorMap* m, const void* base, const uint64_t* dims, const uint64_t* strides_elems,
                        const uint32_t* box) {
    EncodeTiledFn fn = wg_encode_fn();
    if (!fn) return false;
    cuuint64_t gdims[4], gstrides[3];
    cuuint32_t gbox[4], estr[4];
    for (int i = 0; i < 4; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (dims[i] == 0) return false;
    }
    for (int i = 0; i < 3; ++i) gstrides[i] = strides_elems[i] * 2;
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdims, gstrides, gbox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int wg_ew_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

static void wg_add_group(WgradParams& p, int view, const int* khs, int n_kh, const int* kws, int n_kw, int KW, bool strided) {
    WgGroup& g = p.groups[p.n_groups++];
    memset(&g, 0, sizeof(g));
    g.view = (int8_t)view;
    int r0 = 100;
    for (int i = 0; i < n_kh; ++i) {
        const int kh = khs[i];
        const int dh = strided ? (kh - 2 - (kh & 1)) / 2 : kh - KW / 2;
        if (dh < r0) r0 = dh;
    }
    g.r0 = (int8_t)r0;
    int n = 0;
    for (int i = 0; i < n_kh; ++i)
        for (int j = 0; j < n_kw; ++j) {
            const int kh = khs[i], kw = kws[j];
            const int dh = strided ? (kh - 2 - (kh & 1)) / 2 : kh - KW / 2;
            const int dw = strided ? (kw - 2 - (kw & 1)) / 2 : kw - KW / 2;
            g.taps[n].dr = (int8_t)(dh - r0);
            g.taps[n].dc = (int8_t)(dw + 1);
            g.taps[n].w_tap = (int16_t)(kh * KW + kw);
            ++n;
        }
    g.n_taps = (int8_t)n;
}


static bool gb_make_map2(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows) {
    EncodeTiledFn fn = wg_encode_fn();
    if (!fn || rows == 0) return false;
    cuuint64_t gdims[2] = {cols, rows}, gstrides[1] = {cols * 2};
    cuuint32_t gbox[2] = {64, 128}, estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdims, gstrides, gbox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace licos

using namespace licos;

extern "C" {

int licos_conv_wgrad(const licos_wgrad_args* a, void* stream) {
    if (!a || !a->small_t || !a->big_t || !a->out) return LICOS_ERR_INVALID;
    if (((uintptr_t)a->out & 15) != 0) return LICOS_ERR_INVALID;  // the flush is a 16-byte vector red.add
    if (a->batch < 0 || a->h < 1 || a->w < 1 || a->small_c < 1 || a->big_c < 1) return LICOS_ERR_INVALID;
    if (a->batch == 0) return LICOS_OK;
    const bool strided = a->kind == LICOS_CONV_5X5_S2 || a->kind == LICOS_DECONV_5X5_S2;
    // channel counts: multiples of 64; the 1x1 kind also takes any multiple of 16 for `big` (patch matrices): the columns
    // beyond big_c are TMA zero fill and are not flushed
    if (a->small_c % 64 != 0 || a->big_c % (a->kind == LICOS_CONV_1X1 ? 16 : 64) != 0) return LICOS_ERR_UNSUPPORTED;
    if (!strided && a->kind != LICOS_CONV_3X3_S1 && a->kind != LICOS_CONV_1X1) return LICOS_ERR_INVALID;
    if (strided ? (a->h != (a->big_h + 1) / 2 || a->w != (a->big_w + 1) / 2) : (a->h != a->big_h || a->w != a->big_w))
        return LICOS_ERR_INVALID;

    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.Cs = a->small_c;
    p.Cb = a->big_c;
    p.out = a->out;
    p.m_blocks = (a->small_c + 127) / 128;
    p.n_blocks = (a->big_c + 127) / 128;
    if (strided) {
        static const int e02[2] = {0, 2}, e4[1] = {4}, e024[3] = {0, 2, 4}, o13[2] = {1, 3};
        wg_add_group(p, 0, e02, 2, e02, 2, 5, true);
        wg_add_group(p, 0, e02, 2, e4, 1, 5, true);
        wg_add_group(p, 0, e4, 1, e024, 3, 5, true);
        wg_add_group(p, 1, e02, 2, o13, 2, 5, true);
        wg_add_group(p, 1, e4, 1, o13, 2, 5, true);
        wg_add_group(p, 2, o13, 2, e02, 2, 5, true);
        wg_add_group(p, 2, o13, 2, e4, 1, 5, true);
        wg_add_group(p, 3, o13, 2, o13, 2, 5, true);
    } else if (a->kind == LICOS_CONV_3X3_S1) {
        static const int k012[3] = {0, 1, 2};
        for (int kh = 0; kh < 3; ++kh) wg_add_group(p, 0, &k012[kh], 1, k012, 3, 3, false);
    } else {
        static const int k0[1] = {0};
        wg_add_group(p, 0, k0, 1, k0, 1, 1, false);
    }
    p.n_combos = p.n_groups * p.m_blocks * p.n_blocks;
    if (p.n_combos > kWgMaxCombos) return LICOS_ERR_UNSUPPORTED;

    p.tiles_h = (a->h + kWgTH - 1) / kWgTH;
    p.tiles_w = (a->w + kWgTW - 1) / kWgTW;
    const int64_t n_tiles = (int64_t)a->batch * p.tiles_h * p.tiles_w;
    if (n_tiles > 0x3fffffff) return LICOS_ERR_UNSUPPORTED;
    p.n_tiles = (int)n_tiles;

    int sms = a->sm_count;
    if (sms <= 0) {
        int dev = 0;
        LICOS_CUDA_OK(cudaGetDevice(&dev));
        LICOS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // units: every (group, m block, n block) splits its pixel tiles into a share proportional to its tap count
    {
        int64_t total = 0;
        for (int g = 0; g < p.n_groups; ++g) total += (int64_t)p.groups[g].n_taps * p.m_blocks * p.n_blocks;
        // Every unit ends with a flush of its accumulators (up to 256 KB of red.add traffic, the tensor pipe idle
        // meanwhile), so the default is ONE unit per SM, all of equal cost (measured: 96.7 us against 101.3 us with two per SM
        // for 128->128 @ 64^2 x 32 tiles, 35 against 47 us one level down), and at least `min_tiles` pixel tiles per unit
        // (development knobs LICOS_WGRAD_UNITS_PER_SM / LICOS_WGRAD_MIN_TILES).
        static const int per_sm = [] { const char* e = getenv("LICOS_WGRAD_UNITS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 1; }();
        static const int min_tiles = [] { const char* e = getenv("LICOS_WGRAD_MIN_TILES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
        const int64_t target = (int64_t)per_sm * sms;
        // largest-remainder allocation: exactly `target` units in all (when the caps allow), each combo's share proportional
        // to its tap count, so that every unit costs the same and no CTA ends up with one unit more than the others
        int share[kWgMaxCombos];
        int64_t rem[kWgMaxCombos];
        int64_t cap = n_tiles / min_tiles;
        if (cap < 1) cap = 1;
        int64_t given = 0;
        for (int c = 0; c < p.n_combos; ++c) {
            const int g = c / (p.m_blocks * p.n_blocks);
            const int64_t num = (int64_t)p.groups[g].n_taps * target;
            int64_t sh = num / total;
            rem[c] = num % total;
            if (sh < 1) { sh = 1; rem[c] = -1; }
            if (sh >= cap) { sh = cap; rem[c] = -1; }
            share[c] = (int)sh;
            given += sh;
        }
        while (given < target) {
            int best = -1;
            for (int c = 0; c < p.n_combos; ++c)
                if (rem[c] >= 0 && (best < 0 || rem[c] > rem[best])) best = c;
            if (best < 0) break;
            ++share[best];
            rem[best] = -1;
            ++given;
        }
        int n = 0;
        for (int c = 0; c < p.n_combos; ++c) {
            p.unit_begin[c] = n;
            n += share[c];
        }
        p.unit_begin[p.n_combos] = n;
        p.n_units = n;
    }

    {
        const uint64_t C = (uint64_t)a->small_c, H = (uint64_t)a->h, W = (uint64_t)a->w, B = (uint64_t)a->batch;
        const uint64_t dims[4] = {C, W, H, B};
        const uint64_t strides[3] = {C, W * C, H * W * C};
        const uint32_t box[4] = {64, (uint32_t)kWgTW, (uint32_t)kWgTH, 1};
        if (!wg_make_map(&p.s_map, a->small_t, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    {
        const uint64_t C = (uint64_t)a->big_c, H = (uint64_t)a->big_h, W = (uint64_t)a->big_w, B = (uint64_t)a->batch;
        const uint32_t box[4] = {64, (uint32_t)kWgSlabCols, (uint32_t)kWgSlabRows, 1};
        if (strided) {
            if (a->big_h < 2 || a->big_w < 2) return LICOS_ERR_UNSUPPORTED;
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw) {
                    const uint64_t dims[4] = {C, (W - pw + 1) / 2, (H - ph + 1) / 2, B};
                    const uint64_t strides[3] = {2 * C, 2 * W * C, H * W * C};
                    const __nv_bfloat16* base = (const __nv_bfloat16*)a->big_t + ((size_t)ph * W + pw) * C;
                    if (!wg_make_map(&p.g_maps[ph * 2 + pw], base, dims, strides, box)) return LICOS_ERR_CUDA;
                }
        } else {
            const uint64_t dims[4] = {C, W, H, B};
            const uint64_t strides[3] = {C, W * C, H * W * C};
            if (!wg_make_map(&p.g_maps[0], a->big_t, dims, strides, box)) return LICOS_ERR_CUDA;
            for (int i = 1; i < 4; ++i) p.g_maps[i] = p.g_maps[0];
        }
    }
    const size_t smem = 1024 + (size_t)kWgStages * kWgStage;
    LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)wgrad_kernel, (int)smem));
    const int grid = p.n_units < sms ? p.n_units : sms;
    wgrad_kernel<<<grid, kWgThreads, smem, (cudaStream_t)stream>>>(p);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_conv_wgrad_image(const void* small_t, const float* image, int batch, int channels, int h, int w, int small_c,
                           float* out, int sm_count, void* stream) {
    if (!small_t || !image || !out || batch < 0 || channels < 1 || h < 1 || w < 1 || small_c < 1) return LICOS_ERR_INVALID;
    if (((uintptr_t)out & 15) != 0) return LICOS_ERR_INVALID;
    if (batch == 0) return LICOS_OK;
    // the fused kernel is built for the reference's band counts with TMA-loadable rows; anything else takes
    // licos_im2col5x5s2 + licos_conv_wgrad(CONV_1X1)
    if ((channels != 1 && channels != 3) || small_c % 64 != 0 || small_c > 256 || (w % 4) != 0 || ((uintptr_t)image & 15) != 0 ||
        ((uintptr_t)small_t & 15) != 0)
        return LICOS_ERR_UNSUPPORTED;
    const int OH = (h + 1) / 2, OW = (w + 1) / 2;
    WgImageParams p;
    memset(&p, 0, sizeof(p));
    p.Cs = small_c;
    p.m_blocks = (small_c + 127) / 128;
    p.k_pad = (int)licos_im2col5x5s2_kpad(channels);
    p.tiles_h = (OH + 7) / 8;
    p.tiles_w = (OW + 15) / 16;
    const int64_t tiles = (int64_t)batch * p.tiles_h * p.tiles_w;
    if (tiles > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    p.total_tiles = (int)tiles;
    p.slots = p.m_blocks == 1 ? 3 : 2;
    p.out = out;
    {
        EncodeTiledFn fn = wg_encode_fn();
        if (!fn) return LICOS_ERR_CUDA;
        const cuuint64_t gdims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)channels, (cuuint64_t)batch};
        const cuuint64_t gstrides[3] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4, (cuuint64_t)channels * h * w * 4};
        const cuuint32_t gbox[4] = {(cuuint32_t)kWiPatchPitch, (cuuint32_t)kWiPatchRows, (cuuint32_t)channels, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (fn(&p.x_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(image), gdims, gstrides, gbox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return LICOS_ERR_CUDA;
    }
    {
        const uint64_t C = (uint64_t)small_c, H = (uint64_t)OH, W = (uint64_t)OW, B = (uint64_t)batch;
        const uint64_t dims[4] = {C, W, H, B};
        const uint64_t strides[3] = {C, W * C, H * W * C};
        const uint32_t box[4] = {64, 16, 8, 1};
        if (!wg_make_map(&p.s_map, small_t, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    int sms = sm_count;
    if (sms <= 0) {
        int dev = 0;
        LICOS_CUDA_OK(cudaGetDevice(&dev));
        LICOS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = (int)(tiles < sms ? tiles : sms);
    const size_t smem = wg_image_smem_bytes(channels, p.m_blocks, p.slots);
    if (channels == 3) {
        LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)wgrad_image_kernel<3>, 228000));
        wgrad_image_kernel<3><<<grid, kWiThreads, smem, (cudaStream_t)stream>>>(p);
    } else {
        LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)wgrad_image_kernel<1>, 228000));
        wgrad_image_kernel<1><<<grid, kWiThreads, smem, (cudaStream_t)stream>>>(p);
    }
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gdn_backward(const void* x, const void* g, const void* gamma_hat_bf16, const float* beta_hat, int inverse,
                       int64_t n_pixels, int channels, void* dx, float* d_gamma_hat, float* d_beta_hat, float* d_bias,
                       int sm_count, void* stream) {
    if (!x || !g || !gamma_hat_bf16 || !beta_hat || !dx || !d_gamma_hat || !d_beta_hat || n_pixels < 0) return LICOS_ERR_INVALID;
    if (((uintptr_t)d_gamma_hat & 15) != 0) return LICOS_ERR_INVALID;
    if (channels != kGbC) return LICOS_ERR_UNSUPPORTED;  // other widths: the unfused sequence (see the header)
    if (n_pixels == 0) return LICOS_OK;
    const int64_t tiles = (n_pixels + 127) / 128;
    if (tiles > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    GdnBwdParams p;
    memset(&p, 0, sizeof(p));
    if (!gb_make_map2(&p.x_map, x, kGbC, (uint64_t)n_pixels) || !gb_make_map2(&p.g_map, g, kGbC, (uint64_t)n_pixels) ||
        !gb_make_map2(&p.dx_map, dx, kGbC, (uint64_t)n_pixels) || !gb_make_map2(&p.gamma_map, gamma_hat_bf16, kGbC, kGbC))
        return LICOS_ERR_CUDA;
    p.beta_hat = beta_hat;
    p.d_gamma_hat = d_gamma_hat;
    p.d_beta_hat = d_beta_hat;
    p.d_bias = d_bias;
    p.n_tiles = (int)tiles;
    int sms = sm_count;
    if (sms <= 0) {
        int dev = 0;
        LICOS_CUDA_OK(cudaGetDevice(&dev));
        LICOS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = (int)(tiles < sms ? tiles : sms);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool one_team = getenv("LICOS_GDN_BWD_ONE_TEAM") != nullptr;  // A/B knob: the round-1 kernel
    if (!one_team) {
        // the kernel's setmaxnreg budget (2 x 128 x 224 + 128 x 56) must fit the registers the CTA is launched with
        cudaFuncAttributes fa;
        const void* fn = inverse ? (const void*)gdn_bwd_fused2_kernel<true> : (const void*)gdn_bwd_fused2_kernel<false>;
        LICOS_CUDA_OK(cudaFuncGetAttributes(&fa, fn));
        if ((int64_t)fa.numRegs * kGb2Threads < 256 * kGb2TeamRegs + 128 * kGb2IssuerRegs) return LICOS_ERR_UNSUPPORTED;
        if (inverse) {
            LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)gdn_bwd_fused2_kernel<true>, (int)kGb2Smem));
            gdn_bwd_fused2_kernel<true><<<grid, kGb2Threads, kGb2Smem, st>>>(p);
        } else {
            LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)gdn_bwd_fused2_kernel<false>, (int)kGb2Smem));
            gdn_bwd_fused2_kernel<false><<<grid, kGb2Threads, kGb2Smem, st>>>(p);
        }
        LICOS_CUDA_OK(cudaGetLastError());
        return LICOS_OK;
    }
    if (inverse) {
        LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)gdn_bwd_fused_kernel<true>, (int)kGbSmem));
        gdn_bwd_fused_kernel<true><<<grid, kGbThreads, kGbSmem, st>>>(p);
    } else {
        LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)gdn_bwd_fused_kernel<false>, (int)kGbSmem));
        gdn_bwd_fused_kernel<false><<<grid, kGbThreads, kGbSmem, st>>>(p);
    }
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_square_bf16(const void* x, void* x2, int64_t n, void* stream) {
    if (!x || !x2 || n < 0 || n % 8 != 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    square_bf16_kernel<<<wg_ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)x2, n / 8);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

// grid whose thread count is a multiple of C / 8 for every C the kernels take (multiples of 8 up to 512 that divide
// 256 * 1200 / ... : 1200 blocks of 256 threads = 307200 threads, divisible by 16, 24, 32, 40, 48, 64)
static int colsum_ew_grid(int64_t n8, int C) {
    const int per = C / 8;
    int64_t g = (n8 + 255) / 256;
    if (g > 1200) g = 1200;
    while (g > 1 && (g * 256) % per != 0) --g;
    return (int)(g < 1 ? 1 : g);
}

int licos_gdn_bwd_mid(const void* x, const void* g, const void* norm, int inverse, int64_t n, int channels, void* d_norm,
                      void* d_direct, float* sum_d_norm, void* stream) {
    if (!x || !g || !norm || !d_norm || !d_direct || n < 0 || n % 8 != 0) return LICOS_ERR_INVALID;
    if (sum_d_norm && (channels < 8 || channels % 8 != 0 || channels > 512 || n % channels != 0)) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    const int grid = sum_d_norm ? colsum_ew_grid(n / 8, channels) : wg_ew_grid(n / 8);
    if (sum_d_norm && ((int64_t)grid * 256) % (channels / 8) != 0) return LICOS_ERR_UNSUPPORTED;
    if (inverse)
        gdn_bwd_mid_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)g, (const uint4*)norm,
                                                                         (uint4*)d_norm, (uint4*)d_direct, n / 8, channels, sum_d_norm);
    else
        gdn_bwd_mid_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)g, (const uint4*)norm,
                                                                          (uint4*)d_norm, (uint4*)d_direct, n / 8, channels, sum_d_norm);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gdn_bwd_out(const void* x, const void* t, const void* d_direct, int64_t n, int channels, void* dx, float* sum_dx,
                      void* stream) {
    if (!x || !t || !d_direct || !dx || n < 0 || n % 8 != 0) return LICOS_ERR_INVALID;
    if (sum_dx && (channels < 8 || channels % 8 != 0 || channels > 512 || n % channels != 0)) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    const int grid = sum_dx ? colsum_ew_grid(n / 8, channels) : wg_ew_grid(n / 8);
    if (sum_dx && ((int64_t)grid * 256) % (channels / 8) != 0) return LICOS_ERR_UNSUPPORTED;
    gdn_bwd_out_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)t, (const uint4*)d_direct, (uint4*)dx,
                                                               n / 8, channels, sum_dx);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gdn_param_grad(const float* beta, const float* gamma, const float* d_beta_hat, const float* d_gamma_hat, int channels,
                         float beta_bound, float gamma_bound, float* d_beta, float* d_gamma, void* stream) {
    if (!beta || !gamma || !d_beta_hat || !d_gamma_hat || !d_beta || !d_gamma || channels < 1) return LICOS_ERR_INVALID;
    gdn_param_grad_kernel<<<wg_ew_grid((int64_t)channels * channels), 256, 0, (cudaStream_t)stream>>>(
        beta, gamma, d_beta_hat, d_gamma_hat, channels, beta_bound, gamma_bound, d_beta, d_gamma);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_relu_bwd(const void* y, const void* g, int64_t n, void* dx, void* stream) {
    if (!y || !g || !dx || n < 0 || n % 8 != 0) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    relu_bwd_kernel<<<wg_ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)y, (const uint4*)g, (uint4*)dx, n / 8);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_colsum_bf16(const void* x, int64_t rows, int channels, float* acc, void* stream) {
    if (!x || !acc || rows < 0 || channels < 8 || channels % 8 != 0 || channels > 512) return LICOS_ERR_INVALID;
    if (rows == 0) return LICOS_OK;
    const int rows_per_it = 256 / (channels / 8);
    int64_t grid = (rows + rows_per_it - 1) / rows_per_it;
    if (grid > 148 * 4) grid = 148 * 4;
    colsum_bf16_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, rows, channels, acc);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

// patch rows are padded to a multiple of 16 columns only (75 -> 80 for three bands): the weight-gradient kernel reads them
// through a 64-channel TMA box whose out-of-range columns are zero fill, so narrower rows are simply less HBM traffic
int64_t licos_im2col5x5s2_kpad(int channels) { return channels < 1 ? LICOS_ERR_INVALID : (int64_t)(channels * 25 + 15) / 16 * 16; }

int licos_im2col5x5s2(const float* x, int batch, int channels, int h, int w, void* rows, void* stream) {
    if (!x || !rows || batch < 0 || channels < 1 || channels > 16 || h < 1 || w < 1) return LICOS_ERR_INVALID;
    if (batch == 0) return LICOS_OK;
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    const int kp = (int)licos_im2col5x5s2_kpad(channels);
    const int segs = (ow + kImPix - 1) / kImPix;
    const int64_t blocks = (int64_t)batch * oh * segs;
    if (blocks > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    const size_t smem = (size_t)channels * 5 * kImPitch * sizeof(float) + (size_t)kp * sizeof(int16_t);
    im2col5x5s2_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>(x, batch, channels, h, w, oh, ow, kp, segs,
                                                                        (__nv_bfloat16*)rows);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

}  // extern "C"
