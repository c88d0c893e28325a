// Fused GDN / IGDN backward for 128 channels (SURVEY.md 8a row A5 differentiated): one pass over x and g.
//
//   norm = beta + gamma . x^2           GEMM 1   [128 px x 128 i] . gamma^T        (A = x^2 tile, K-major; B = gamma, K-major)
//   d_norm = g * dy/dnorm, d_direct = g * dy/dx|norm                               (threads; one pixel row each)
//   t = gamma^T . d_norm                GEMM 2   A = d_norm tile, K-major; B = the SAME resident gamma read MN-major
//   gamma_hat.grad += d_norm^T . x^2    GEMM 3   both tiles read MN-major (contraction over the 128 pixels), accumulator
//                                                resident in TMEM for the whole kernel
//   beta_hat.grad  += d_norm^T . 1      GEMM 3b  against a constant "ones" operand (N = 16)
//   dx = d_direct + 2 x t                                                          (threads) -> bf16 -> TMA store
//   bias.grad      += dx^T . 1          GEMM 4
//
// The unfused sequence (square, 1x1 layer, mid, 1x1 layer, out, wgrad, column sums) moves 14 tensor-sized streams through
// HBM; this kernel moves three (x, g in; dx out).  Pixels are a flat [P][128] matrix: a tile is 128 consecutive pixels,
// the ragged tail is TMA zero fill / store clipping (x = g = 0 contributes nothing).  One CTA per SM, 5 warps: four
// pixel-row warps (TMEM lane quarters) whose first lane group also issues the MMAs, one TMA producer warp that keeps the
// next tile's x and g in flight.
// Measured (ncu source page): ~5.9 k warp instructions per tile and warp, i.e. the kernel is bound by the issue rate of the
// pixel-row code (bf16 unpack / pack, scalar fp32 math), not by latency: a two-team variant with a dedicated issuer warp
// (two warps per scheduler) was built and measured at the same 165-170 us for 32 x 128 x 128 pixels, so it was dropped.
// Steps 2 and 3 use the packed fp32x2 arithmetic of epilogue.cuh (FMUL2 / FADD2 / FFMA2).
#pragma once

#include "common.cuh"
#include "epilogue.cuh"

namespace licos {

constexpr int kGbC = 128;
constexpr uint32_t kGbChunk = 128u * 128u;      // [128 rows][64 bf16]
constexpr uint32_t kGbTile = 2u * kGbChunk;     // [128 rows][128 bf16]
constexpr int kGbThreads = 160;
// shared memory: gamma (32 KB), ones (16 KB), x/x^2 double buffer (64 KB), g/dx double buffer (64 KB), d_norm (32 KB)
constexpr uint32_t kGbSmem = 1024u + kGbTile + kGbChunk + 2u * kGbTile + 2u * kGbTile + kGbTile;

struct GdnBwdParams {
    CUtensorMap x_map, g_map, dx_map, gamma_map;
    const float* beta_hat;
    float* d_gamma_hat;  // [128][128], accumulated
    float* d_beta_hat;   // [128], accumulated
    float* d_bias;       // [128], accumulated, may be NULL
    int n_tiles;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ uint64_t gb_desc_mn(uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void load_row32(const uint8_t* tile_base, int row, int col32, uint32_t (&pk)[16]) {
    const uint8_t* atom = tile_base + (size_t)((col32 * 32) / 64) * (128 * 128);
    const uint32_t chunk0 = ((col32 * 32) % 64) / 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(atom + sw128_offset(row, chunk0 + q));
        pk[4 * q] = v.x; pk[4 * q + 1] = v.y; pk[4 * q + 2] = v.z; pk[4 * q + 3] = v.w;
    }
}

template <bool INVERSE>
__global__ void __launch_bounds__(kGbThreads, 1) gdn_bwd_fused_kernel(const __grid_constant__ GdnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[2], empty[2], g_full, mma_bar, g4_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float beta_s[kGbC];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    uint8_t* gamma_s = smem;
    uint8_t* ones_s = gamma_s + kGbTile;
    uint8_t* xbuf = ones_s + kGbChunk;     // 2 tiles
    uint8_t* gbuf = xbuf + 2 * kGbTile;    // 2 tiles
    uint8_t* dn_s = gbuf + 2 * kGbTile;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&g_full, 1);
        mbar_init(&mma_bar, 1);
        mbar_init(&g4_bar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < kGbC) beta_s[threadIdx.x] = p.beta_hat[threadIdx.x];
    // constant operand: ones[pixel][channel 0] = 1, everything else 0 (16 channels of it are read as N = 16)
    for (uint32_t i = threadIdx.x; i < kGbChunk / 16; i += kGbThreads) reinterpret_cast<uint4*>(ones_s)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < 128) *reinterpret_cast<uint16_t*>(ones_s + sw128_offset(threadIdx.x, 0)) = 0x3F80;  // bf16 1.0
    fence_proxy_async();
    if (warp == 4) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 4) {
        if (lane == 0) {
            // ===================== producer =====================
            tma_prefetch_desc(&p.x_map);
            tma_prefetch_desc(&p.g_map);
            mbar_arrive_expect_tx(&g_full, kGbTile);
            for (int gc = 0; gc < 2; ++gc) tma_load_2d(gamma_s + gc * kGbChunk, &p.gamma_map, &g_full, gc * 64, 0);
            for (int k = 0; k < my_tiles; ++k) {
                const int b = k & 1;
                const int row0 = ((int)blockIdx.x + k * (int)gridDim.x) * 128;
                mbar_wait(&empty[b], ((uint32_t)(k >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&full[b], 2 * kGbTile);
                for (int c = 0; c < 2; ++c) {
                    tma_load_2d(xbuf + b * kGbTile + c * kGbChunk, &p.x_map, &full[b], c * 64, row0);
                    tma_load_2d(gbuf + b * kGbTile + c * kGbChunk, &p.g_map, &full[b], c * 64, row0);
                }
            }
        }
    } else {
        // ===================== pixel-row team (threads 0..127 == TMEM lanes) =====================
        const int et = threadIdx.x;
        const uint32_t lane_sel = ((uint32_t)warp * 32u) << 16;
        const bool leader = et == 0;
        const uint32_t t_norm = tmem_base + lane_sel;  // columns [0,128): norm, then t
        const uint32_t idesc_kk = umma_idesc_bf16(128, 128);
        const uint32_t idesc_kmn = idesc_kk | (1u << 16);
        const uint32_t idesc_mnmn = idesc_kk | (1u << 15) | (1u << 16);
        const uint32_t idesc_ones = umma_idesc_bf16(128, 16) | (1u << 15) | (1u << 16);
        const uint64_t mn_hi = gb_desc_mn(kGbChunk);
        const uint32_t gamma16 = smem_u32(gamma_s) >> 4, ones16 = smem_u32(ones_s) >> 4, dn16 = smem_u32(dn_s) >> 4;
        uint32_t n_mma = 0;
        mbar_wait(&g_full, 0);
        for (int k = 0; k < my_tiles; ++k) {
            const int b = k & 1;
            uint8_t* xt = xbuf + b * kGbTile;
            uint8_t* gt = gbuf + b * kGbTile;
            const uint32_t x16 = smem_u32(xt) >> 4, g16 = smem_u32(gt) >> 4;
            if (leader && k > 0) {
                // the other buffer pair (tile k-1) is free once its dx store has been read out and GEMM 4 has consumed it
                tma_store_wait_read();
                mbar_wait(&g4_bar, (uint32_t)(k - 1) & 1u);
                mbar_arrive(&empty[b ^ 1]);
            }
            mbar_wait(&full[b], (uint32_t)(k >> 1) & 1u);

            // ---- step 1: x -> registers (packed bf16), x^2 -> in place = A operand of GEMM 1, B operand of GEMM 3 ----
            uint32_t xs[64];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t pk[16], sq[16];
                load_row32(xt, et, cc, pk);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    xs[cc * 16 + j] = pk[j];
                    const uint64_t v = bf16x2_to_f2(pk[j]);
                    sq[j] = f2_to_bf16x2(f2_mul(v, v));
                }
                store_row32(xt, et, cc, sq);
            }
            fence_proxy_async();
            tc_fence_before();
            named_bar_sync(1, 128);
            if (warp == 0) {
                tc_fence_after();
                if (lane == 0) {
                    issue_gamma_gemm<8>(tmem_base, x16, gamma16, 128, idesc_kk);
                    umma_commit(&mma_bar);
                }
                __syncwarp();
            }
            mbar_wait(&mma_bar, n_mma & 1u);
            ++n_mma;
            tc_fence_after();

            // ---- step 2: d_norm -> smem (A of GEMM 2, 3, 3b), d_direct -> registers ----
            uint32_t dd[64];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float v[32];
                tmem_ld32(t_norm + cc * 32, v);
                uint32_t gp[16], dn[16];
                load_row32(gt, et, cc, gp);
                tmem_ld_wait();
                const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(beta_s + cc * 32);
                const uint64_t half = INVERSE ? f2_pack(0.5f, 0.5f) : f2_pack(-0.5f, -0.5f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const ulonglong2 bq = b2[q];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = 2 * q + h;
                        const uint64_t d = f2_add(f2_pack(v[2 * j], v[2 * j + 1]), h ? bq.y : bq.x);
                        float d0, d1;
                        f2_unpack(d, d0, d1);
                        const uint64_t r = f2_pack(fast_rsqrt(d0), fast_rsqrt(d1));
                        const uint64_t g2 = bf16x2_to_f2(gp[j]);
                        const uint64_t gx = f2_mul(g2, bf16x2_to_f2(xs[cc * 16 + j]));
                        uint64_t ddv, dnv;
                        if (INVERSE) {  // y = x sqrt(d): dy/dx = sqrt(d) = d r, dy/dd = x r / 2
                            ddv = f2_mul(g2, f2_mul(d, r));
                            dnv = f2_mul(gx, f2_mul(r, half));
                        } else {        // y = x rsqrt(d): dy/dx = r, dy/dd = -x r^3 / 2
                            ddv = f2_mul(g2, r);
                            dnv = f2_mul(gx, f2_mul(f2_mul(r, r), f2_mul(r, half)));
                        }
                        dd[cc * 16 + j] = f2_to_bf16x2(ddv);
                        dn[j] = f2_to_bf16x2(dnv);
                    }
                }
                store_row32(dn_s, et, cc, dn);
            }
            fence_proxy_async();
            tc_fence_before();
            named_bar_sync(1, 128);
            if (warp == 0) {
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t k_hi = umma_desc_sw128(0);
                    // GEMM 2: t[px][i] = sum_o d_norm[px][o] gamma[o][i]   (B = gamma read MN-major: N = i contiguous, K = o rows)
#pragma unroll
                    for (uint32_t ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem_base, k_hi | (uint64_t)(dn16 + (ks >> 2) * 1024 + (ks & 3) * 2), mn_hi | (uint64_t)(gamma16 + ks * 128),
                                  idesc_kmn, (uint32_t)(ks > 0));
                    // GEMM 3: gamma_hat.grad[o][i] += sum_px d_norm[px][o] x2[px][i]   (both MN-major, K = pixels)
#pragma unroll
                    for (uint32_t ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem_base + 128, mn_hi | (uint64_t)(dn16 + ks * 128), mn_hi | (uint64_t)(x16 + ks * 128), idesc_mnmn,
                                  (uint32_t)(k > 0 || ks > 0));
                    // GEMM 3b: beta_hat.grad[o] += sum_px d_norm[px][o]
#pragma unroll
                    for (uint32_t ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem_base + 256, mn_hi | (uint64_t)(dn16 + ks * 128), mn_hi | (uint64_t)(ones16 + ks * 128), idesc_ones,
                                  (uint32_t)(k > 0 || ks > 0));
                    umma_commit(&mma_bar);
                }
                __syncwarp();
            }
            mbar_wait(&mma_bar, n_mma & 1u);
            ++n_mma;
            tc_fence_after();

            // ---- step 3: dx = d_direct + 2 x t -> the g buffer (staging for the TMA store, A of GEMM 4) ----
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float v[32];
                tmem_ld32(t_norm + cc * 32, v);
                tmem_ld_wait();
                uint32_t out[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint64_t x2 = bf16x2_to_f2(xs[cc * 16 + j]);
                    out[j] = f2_to_bf16x2(f2_fma(f2_add(x2, x2), f2_pack(v[2 * j], v[2 * j + 1]), bf16x2_to_f2(dd[cc * 16 + j])));
                }
                store_row32(gt, et, cc, out);
            }
            fence_proxy_async();
            tc_fence_before();
            named_bar_sync(1, 128);
            if (warp == 0) {
                tc_fence_after();
                if (lane == 0) {  // the thread that commits the bulk group is the one that later waits on it
                    const int row0 = ((int)blockIdx.x + k * (int)gridDim.x) * 128;
                    tma_store_2d(&p.dx_map, gt, 0, row0);
                    tma_store_2d(&p.dx_map, gt + kGbChunk, 64, row0);
                    tma_store_commit();
                    // GEMM 4: bias.grad[c] += sum_px dx[px][c]
#pragma unroll
                    for (uint32_t ks = 0; ks < 8; ++ks)
                        umma_bf16(tmem_base + 272, mn_hi | (uint64_t)(g16 + ks * 128), mn_hi | (uint64_t)(ones16 + ks * 128), idesc_ones,
                                  (uint32_t)(k > 0 || ks > 0));
                    umma_commit(&g4_bar);
                }
                __syncwarp();
            }
        }
        // ---- flush the resident accumulators ----
        if (my_tiles > 0) {
            mbar_wait(&g4_bar, (uint32_t)(my_tiles - 1) & 1u);
            tc_fence_after();
            float* o = p.d_gamma_hat + (size_t)et * kGbC;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float v[32];
                tmem_ld32(tmem_base + lane_sel + 128 + cc * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + cc * 32 + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                                 "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                 : "memory");
            }
            float s[16];
            tmem_ld16(tmem_base + lane_sel + 256, s);
            tmem_ld_wait();
            atomicAdd(p.d_beta_hat + et, s[0]);
            if (p.d_bias) {
                tmem_ld16(tmem_base + lane_sel + 272, s);
                tmem_ld_wait();
                atomicAdd(p.d_bias + et, s[0]);
            }
        }
        if (leader) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace licos
