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

#include <stdio.h>

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

// pulls a box into L2 only (no shared memory, no barrier): issued one tile ahead of the real load
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
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

// 16 channels (32 B) of one pixel row: two 16-byte chunks
__device__ __forceinline__ void load_row16(const uint8_t* tile_base, int row, int col16, uint32_t (&pk)[8]) {
    const uint8_t* atom = tile_base + (size_t)(col16 >> 2) * (128 * 128);
    const uint32_t chunk0 = (uint32_t)(col16 & 3) * 2u;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(atom + sw128_offset(row, chunk0 + q));
        pk[4 * q] = v.x; pk[4 * q + 1] = v.y; pk[4 * q + 2] = v.z; pk[4 * q + 3] = v.w;
    }
}
__device__ __forceinline__ void store_row16(uint8_t* tile_base, int row, int col16, const uint32_t (&pk)[8]) {
    uint8_t* atom = tile_base + (size_t)(col16 >> 2) * (128 * 128);
    const uint32_t chunk0 = (uint32_t)(col16 & 3) * 2u;
#pragma unroll
    for (int q = 0; q < 2; ++q)
        *reinterpret_cast<uint4*>(atom + sw128_offset(row, chunk0 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
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
        const uint64_t ones_hi = mn_hi & ~((uint64_t)0x3FFF << 32);  // SBO = 0: every 8-pixel group reads the same 1 KB
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
                        umma_bf16(tmem_base + 256, mn_hi | (uint64_t)(dn16 + ks * 128), ones_hi | (uint64_t)ones16, idesc_ones,
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
                        umma_bf16(tmem_base + 272, mn_hi | (uint64_t)(g16 + ks * 128), ones_hi | (uint64_t)ones16, idesc_ones,
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

// ----------------------------------------------------------------------------------------------
// Two-team variant (the one the launcher uses).  What the measurements on the one-team kernel above said (ncu + clock64
// probes): issue slots 22 % busy, ~10 k cycles per tile against ~2 k instructions per warp -- one warp per scheduler that
// sits out every TMEM round trip and every MMA batch; and the column-sum MMAs against the ones operand (M128 N16, A
// MN-major) cost ~107 cycles EACH, as much as a full 128 x 128 one (GEMM 2: ~100, GEMM 3 with both operands MN-major: ~140).
// So here:
//   * two teams of four pixel-row warps work on alternate tiles; a dedicated warp issues every MMA and every TMA load
//     (one issuing thread = one well-defined order for the gamma_hat.grad accumulator both teams add into);
//   * twelve warps: a CTA of nine would cap every thread at 168 registers (three warps on one scheduler's 16 K file), so the
//     issuer's warpgroup (warp 8 works, 9-11 idle) hands its registers over with `setmaxnreg` (224 per pixel-row thread);
//   * TMEM loads run one 32-column piece ahead of the arithmetic;
//   * beta_hat.grad and bias.grad (column sums of d_norm and dx) are summed on the CUDA cores from the tiles the threads just
//     wrote (conflict-free 16-byte shared loads; the partial sums stay in registers until the kernel ends) while the team would otherwise wait for GEMM 2 /
//     the TMA store -- 16 MMAs per tile less on the tensor pipe, no fourth round trip;
//   * a team's next tile is pulled into L2 a whole tile ahead (cp.async.bulk.prefetch), so the real load, which can only be
//     issued once GEMM 3 has released the buffers, is an L2 hit (~1 k cycles instead of ~5 k).
// Shared memory (230 400 B): gamma 32 KB and per team three 32 KB tiles, each rewritten in place by the thread that owns
// the pixel row: P (x -> x^2), BG (g -> d_norm), D (d_direct -> dx).  x itself lives in registers (64 packed words).
// ----------------------------------------------------------------------------------------------
constexpr int kGb2Threads = 384;
constexpr int kGb2TeamRegs = 224, kGb2IssuerRegs = 56;
constexpr uint32_t kGb2Smem = 1024u + kGbTile + 6u * kGbTile;

// acc += this thread's share of the column sums of a [128 px][128 ch] bf16 tile (two swizzled 64-channel atoms): the
// 16-byte chunk et % 16 (8 channels) of rows (et / 16) * 16 .. + 15.  The share is the same in every tile, so the partial
// sums stay in registers for the whole kernel and meet the other row groups only once, at the end.
__device__ __forceinline__ void gb_colsum_acc(const uint8_t* tile, int et, uint64_t (&acc)[4]) {
    const int q = et & 15, r0 = (et >> 4) * 16;
    const uint8_t* atom = tile + (size_t)(q >> 3) * kGbChunk;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const uint4 v = *reinterpret_cast<const uint4*>(atom + sw128_offset(r0 + r, (uint32_t)(q & 7)));
        acc[0] = f2_add(acc[0], bf16x2_to_f2(v.x));
        acc[1] = f2_add(acc[1], bf16x2_to_f2(v.y));
        acc[2] = f2_add(acc[2], bf16x2_to_f2(v.z));
        acc[3] = f2_add(acc[3], bf16x2_to_f2(v.w));
    }
}
// End of the kernel: lanes l and l ^ 16 hold the same eight channels (two row groups) -> one shuffle; lanes 0..15 then write
// the warp's 128 partial sums to scratch[warp][128] (fp32), which the caller adds up over the eight warps.
__device__ __forceinline__ void gb_colsum_stage(int lane, const uint64_t (&acc)[4], float* scratch_warp) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo, hi;
        f2_unpack(acc[i], lo, hi);
        lo += __shfl_xor_sync(0xffffffffu, lo, 16);
        hi += __shfl_xor_sync(0xffffffffu, hi, 16);
        if (lane < 16) {
            scratch_warp[lane * 8 + 2 * i] = lo;
            scratch_warp[lane * 8 + 2 * i + 1] = hi;
        }
    }
}

template <bool INVERSE>
__global__ void __launch_bounds__(kGb2Threads, 1) gdn_bwd_fused2_kernel(const __grid_constant__ GdnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    // per team: x / g landed, rows written after step 1 / 2 / 3 (128 arrivals), GEMM 1 / GEMM 2 / GEMM 3 done, d_norm sums read
    __shared__ uint64_t xfull[2], gfull[2], rows[2][3], mma1[2], mma2[2], free2[2], sums_done[2], g_full, done_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float beta_s[kGbC];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    uint8_t* gamma_s = smem;
    uint8_t* tiles_s = gamma_s + kGbTile;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int t = 0; t < 2; ++t) {
            mbar_init(&xfull[t], 1); mbar_init(&gfull[t], 1);
            for (int i = 0; i < 3; ++i) mbar_init(&rows[t][i], 128);
            mbar_init(&mma1[t], 1); mbar_init(&mma2[t], 1); mbar_init(&free2[t], 1);
            mbar_init(&sums_done[t], 128);
        }
        mbar_init(&g_full, 1);
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < kGbC) {
        beta_s[threadIdx.x] = p.beta_hat[threadIdx.x];
    }
    if (warp == 8) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    constexpr uint32_t kColDg = 256;  // TMEM columns: [0,128) team 0 norm / t, [128,256) team 1, [256,384) gamma_hat.grad

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kGb2IssuerRegs));
        // ===================== MMA + TMA issuer: serves both teams in arrival order =====================
        if (warp == 8 && lane == 0) {
            tma_prefetch_desc(&p.gamma_map);
            mbar_arrive_expect_tx(&g_full, kGbTile);
            for (int gc = 0; gc < 2; ++gc) tma_load_2d(gamma_s + gc * kGbChunk, &p.gamma_map, &g_full, gc * 64, 0);
            const uint32_t idesc_kk = umma_idesc_bf16(128, 128);
            const uint32_t idesc_kmn = idesc_kk | (1u << 16);
            const uint32_t idesc_mnmn = idesc_kk | (1u << 15) | (1u << 16);
            const uint64_t mn_hi = gb_desc_mn(kGbChunk);
            const uint64_t k_hi = umma_desc_sw128(0);
            const uint32_t gamma16 = smem_u32(gamma_s) >> 4, tiles16 = smem_u32(tiles_s) >> 4;
            const int n0 = (my_tiles + 1) / 2, n1 = my_tiles / 2;
            int j0 = 0, j1 = 0, s0 = 0, s1 = 0;  // per team: tile index, next request (0: GEMM 1, 1: GEMM 2 + 3)
            int l0 = 0, l1 = 0;                  // per team: next tile to load
            uint32_t any3 = 0;
            tma_prefetch_desc(&p.x_map);
            tma_prefetch_desc(&p.g_map);
            auto load_tile = [&](int t, int jj) {  // x -> P, g -> BG of team t (x first: step 1 needs only x)
                const int row0 = ((int)blockIdx.x + (2 * jj + t) * (int)gridDim.x) * 128;
                uint8_t* tb = tiles_s + (size_t)t * 3 * kGbTile;
                mbar_arrive_expect_tx(&xfull[t], kGbTile);
                for (int c = 0; c < 2; ++c) tma_load_2d(tb + c * kGbChunk, &p.x_map, &xfull[t], c * 64, row0);
                mbar_arrive_expect_tx(&gfull[t], kGbTile);
                for (int c = 0; c < 2; ++c) tma_load_2d(tb + 2 * kGbTile + c * kGbChunk, &p.g_map, &gfull[t], c * 64, row0);
                if (jj + 1 < (t ? n1 : n0)) {  // the team's following tile -> L2
                    const int row1 = row0 + 2 * (int)gridDim.x * 128;
                    for (int c = 0; c < 2; ++c) {
                        tma_prefetch_l2_2d(&p.x_map, c * 64, row1);
                        tma_prefetch_l2_2d(&p.g_map, c * 64, row1);
                    }
                }
            };
            if (n0 > 0) { load_tile(0, 0); l0 = 1; }
            if (n1 > 0) { load_tile(1, 0); l1 = 1; }
            mbar_wait(&g_full, 0);
            // (rolled loops on purpose: this warp runs on 56 registers, and issuing an MMA blocks for ~100 cycles anyway)
            while (j0 < n0 || j1 < n1) {
                bool served = false;
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    const int jt = t ? j1 : j0, st = t ? s1 : s0, lt = t ? l1 : l0;
                    // the next tile's load: once GEMM 3 has read P and BG and the team has read BG for its column sums
                    if (lt < (t ? n1 : n0) && mbar_test_wait_addr(smem_u32(&free2[t]), (uint32_t)(lt - 1) & 1u) &&
                        mbar_test_wait_addr(smem_u32(&sums_done[t]), (uint32_t)(lt - 1) & 1u)) {
                        load_tile(t, lt);
                        if (t) ++l1; else ++l0;
                        served = true;
                    }
                    if (jt >= (t ? n1 : n0)) continue;
                    if (!mbar_test_wait_addr(smem_u32(&rows[t][st]), (uint32_t)jt & 1u)) continue;
                    tc_fence_after();
                    const uint32_t b16 = tiles16 + (uint32_t)t * ((3u * kGbTile) >> 4);
                    const uint32_t q16 = b16;                             // P: x^2
                    const uint32_t dn16 = b16 + ((2u * kGbTile) >> 4);    // BG: d_norm
                    const uint32_t d_t = tmem_base + (uint32_t)t * 128u;
                    if (st == 0) {
#pragma unroll 1
                        for (uint32_t ks = 0; ks < 8; ++ks)  // norm = gamma . x^2   (both K-major)
                            umma_bf16(d_t, k_hi | (uint64_t)(q16 + (ks >> 2) * 1024 + (ks & 3) * 2),
                                      k_hi | (uint64_t)(gamma16 + (ks >> 2) * 1024 + (ks & 3) * 2), idesc_kk, ks);
                        umma_commit(&mma1[t]);
                        if (t) s1 = 1; else s0 = 1;
                    } else {
#pragma unroll 1
                        for (uint32_t ks = 0; ks < 8; ++ks)  // t = gamma^T . d_norm   (B = gamma read MN-major)
                            umma_bf16(d_t, k_hi | (uint64_t)(dn16 + (ks >> 2) * 1024 + (ks & 3) * 2), mn_hi | (uint64_t)(gamma16 + ks * 128),
                                      idesc_kmn, ks);
                        umma_commit(&mma2[t]);  // the team goes on with t; the accumulation below only releases the buffers
#pragma unroll 1
                        for (uint32_t ks = 0; ks < 8; ++ks)  // gamma_hat.grad += d_norm^T . x^2   (both MN-major, K = pixels)
                            umma_bf16(tmem_base + kColDg, mn_hi | (uint64_t)(dn16 + ks * 128), mn_hi | (uint64_t)(q16 + ks * 128), idesc_mnmn,
                                      any3 | ks);
                        umma_commit(&free2[t]);
                        any3 = 1;
                        if (t) { s1 = 0; ++j1; } else { s0 = 0; ++j0; }
                    }
                    served = true;
                }
                if (!served) __nanosleep(32);  // leave the scheduler's issue slots to the two pixel-row warps it also hosts
            }
            umma_commit(&done_bar);
        }
    } else {
        // ===================== pixel-row teams (a team's 128 threads == the TMEM lanes) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kGb2TeamRegs));
        const int team = warp >> 2;
        const int et = (int)threadIdx.x & 127;
        const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
        const bool leader = et == 0;
        const uint32_t t_acc = tmem_base + lane_sel + (uint32_t)team * 128u;
        uint8_t* qt = tiles_s + (size_t)team * 3 * kGbTile;  // P: x -> x^2 in place (the owning thread keeps x in registers)
        uint8_t* xt = qt + kGbTile;                          // D: d_direct -> dx in place
        uint8_t* bg = qt + 2 * kGbTile;                      // BG: g -> d_norm in place
        const int n_j = (my_tiles - team + 1) / 2;
        if (leader) tma_prefetch_desc(&p.dx_map);
        uint64_t acc_beta[4] = {0ull, 0ull, 0ull, 0ull}, acc_bias[4] = {0ull, 0ull, 0ull, 0ull};
        for (int j = 0; j < n_j; ++j) {
            const uint32_t par = (uint32_t)j & 1u;
            mbar_wait(&xfull[team], par);

            // ---- step 1: x -> registers (packed bf16), x^2 -> P in place = A of GEMM 1, B of GEMM 3 ----
            uint32_t xs[64];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t pk[16], sq[16];
                load_row32(qt, et, cc, pk);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    xs[cc * 16 + i] = pk[i];
                    const uint64_t v = bf16x2_to_f2(pk[i]);
                    sq[i] = f2_to_bf16x2(f2_mul(v, v));
                }
                store_row32(qt, et, cc, sq);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&rows[team][0]);
            mbar_wait(&mma1[team], par);
            tc_fence_after();
            // D still holds the previous tile's dx until its TMA store has read it: the leader waited for that before it
            // arrived on rows[0] above, which mma1 follows.
            mbar_wait(&gfull[team], par);

            // ---- step 2: d_norm -> BG in place of g (A of GEMM 2 and 3), d_direct -> D ----
            // (the TMEM load of the next 32 columns is in flight while these 32 are worked on)
            auto step2 = [&](int cc, const float (&v)[32]) {
                uint32_t gp[16], dn[16], dd[16];
                load_row32(bg, et, cc, gp);
                const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(beta_s + cc * 32);
                const uint64_t half = INVERSE ? f2_pack(0.5f, 0.5f) : f2_pack(-0.5f, -0.5f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const ulonglong2 bq = b2[q];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = 2 * q + h;
                        const uint64_t d = f2_add(f2_pack(v[2 * i], v[2 * i + 1]), h ? bq.y : bq.x);
                        float d0, d1;
                        f2_unpack(d, d0, d1);
                        const uint64_t r = f2_pack(fast_rsqrt(d0), fast_rsqrt(d1));
                        const uint64_t g2 = bf16x2_to_f2(gp[i]);
                        const uint64_t gx = f2_mul(g2, bf16x2_to_f2(xs[cc * 16 + i]));
                        uint64_t ddv, dnv;
                        if (INVERSE) {  // y = x sqrt(d): dy/dx = sqrt(d) = d r, dy/dd = x r / 2
                            ddv = f2_mul(g2, f2_mul(d, r));
                            dnv = f2_mul(gx, f2_mul(r, half));
                        } else {        // y = x rsqrt(d): dy/dx = r, dy/dd = -x r^3 / 2
                            ddv = f2_mul(g2, r);
                            dnv = f2_mul(gx, f2_mul(f2_mul(r, r), f2_mul(r, half)));
                        }
                        dd[i] = f2_to_bf16x2(ddv);
                        dn[i] = f2_to_bf16x2(dnv);
                    }
                }
                store_row32(bg, et, cc, dn);
                store_row32(xt, et, cc, dd);
            };
            {
                float va[32], vb[32];
                tmem_ld32(t_acc, va);
                tmem_ld_wait();
                tmem_ld32(t_acc + 32, vb);
                step2(0, va);
                tmem_ld_wait();
                tmem_ld32(t_acc + 64, va);
                step2(1, vb);
                tmem_ld_wait();
                tmem_ld32(t_acc + 96, vb);
                step2(2, va);
                tmem_ld_wait();
                step2(3, vb);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&rows[team][1]);
            // beta_hat.grad += column sums of d_norm, while GEMM 2 runs (every row of the tile is in place once rows[1] completes)
            mbar_wait(&rows[team][1], par);
            gb_colsum_acc(bg, et, acc_beta);
            mbar_arrive(&sums_done[team]);
            mbar_wait(&mma2[team], par);
            tc_fence_after();

            // ---- step 3: dx = d_direct + 2 x t -> D in place (staging of the TMA store) ----
            auto step3 = [&](int cc, const float (&v)[32]) {
                uint32_t dd[16], out[16];
                load_row32(xt, et, cc, dd);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint64_t x2 = bf16x2_to_f2(xs[cc * 16 + i]);
                    out[i] = f2_to_bf16x2(f2_fma(f2_add(x2, x2), f2_pack(v[2 * i], v[2 * i + 1]), bf16x2_to_f2(dd[i])));
                }
                store_row32(xt, et, cc, out);
            };
            {
                float va[32], vb[32];
                tmem_ld32(t_acc, va);
                tmem_ld_wait();
                tmem_ld32(t_acc + 32, vb);
                step3(0, va);
                tmem_ld_wait();
                tmem_ld32(t_acc + 64, va);
                step3(1, vb);
                tmem_ld_wait();
                tmem_ld32(t_acc + 96, vb);
                step3(2, va);
                tmem_ld_wait();
                step3(3, vb);
            }
            fence_proxy_async();
            tc_fence_before();  // the next tile's GEMM 1 overwrites the columns just read
            mbar_arrive(&rows[team][2]);
            mbar_wait(&rows[team][2], par);  // every row of dx is in place (and fenced for the async proxy)
            if (leader) {
                const int row0 = ((int)blockIdx.x + (2 * j + team) * (int)gridDim.x) * 128;
                tma_store_2d(&p.dx_map, xt, 0, row0);
                tma_store_2d(&p.dx_map, xt + kGbChunk, 64, row0);
                tma_store_commit();
            }
            gb_colsum_acc(xt, et, acc_bias);  // bias.grad += column sums of dx
            if (leader) tma_store_wait_read();  // before this thread's next arrival on rows[0]: see step 2
        }
        // ---- flush: gamma_hat.grad from TMEM (team 0 columns 0..63, team 1 the rest), the two sum vectors from shared memory ----
        mbar_wait(&done_bar, 0);
        tc_fence_after();
        float* o = p.d_gamma_hat + (size_t)et * kGbC + team * 64;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            float v[32];
            tmem_ld32(tmem_base + lane_sel + kColDg + team * 64 + cc * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + cc * 32 + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                             "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                             : "memory");
        }
        // the column sums: every MMA and TMA load has finished (done_bar), so team 0's P tile is free to serve as scratch
        // ([tensor][warp][128] fp32 = 8 KB); the TMA stores still in flight read the D tiles only
        named_bar_sync(1, 256);
        float* scratch = reinterpret_cast<float*>(tiles_s);
        gb_colsum_stage(lane, acc_beta, scratch + (size_t)warp * kGbC);
        gb_colsum_stage(lane, acc_bias, scratch + (size_t)(8 + warp) * kGbC);
        named_bar_sync(1, 256);
        if (team == 0 || p.d_bias) {
            const float* src = scratch + (size_t)(team * 8) * kGbC + et;
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += src[w * kGbC];
            atomicAdd((team == 0 ? p.d_beta_hat : p.d_bias) + et, sum);
        }
        if (leader) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace licos
