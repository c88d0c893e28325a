// Generic fused-im2col first layer (any C_in with 25 C_in <= 128, any width): the fallback of conv_first2.cuh for band
// counts and row pitches its TMA patch loader does not cover.
//
//   conv_first_kernel     g_a[0]: Conv2d(C_in <= 5 -> N, 5x5, s2) + GDN/ReLU, fp32 NCHW in, bf16 NHWC out.
//                         K = 25*C_in <= 128.  The im2col tile is built in shared memory by the CTA itself from
//                         a staged input patch (the fp32 -> bf16 / NCHW -> NHWC conversion is fused into it),
//                         then one short tcgen05 GEMM and the same fused GDN epilogue as the main engine.
//
// A plain (not warp-specialised) loop sized so that two CTAs share an SM and overlap each other.
#pragma once

#include "common.cuh"

namespace licos {

constexpr int kEdgeThreads = 256;

// ------------------------------------------------------------------------------------------------
// first layer
// ------------------------------------------------------------------------------------------------
struct FirstParams {
    CUtensorMap w_map;    // [N][k_pad] bf16, box (64, N)
    CUtensorMap g_map;    // gamma [N][N] bf16, box (64, N)
    CUtensorMap out_map;  // NHWC bf16 (N, OW, OH, B), box (64, 16, 8, 1)
    const float* x;
    const float* bias;
    const float* beta;
    int B, C, H, W, OH, OW;
    int N, K, k_pad;
    int tiles_h, tiles_w, total_tiles;
    uint32_t tmem_cols;
};

constexpr int kPatchRows = 2 * 8 + 3;    // 19 input rows feed 8 output rows
constexpr int kPatchCols = 2 * 16 + 3;   // 35 input columns feed 16 output columns
constexpr int kPatchPitch = 36;

template <int EPI>
__global__ void __launch_bounds__(kEdgeThreads, 2) conv_first_kernel(const __grid_constant__ FirstParams p) {
    constexpr bool kGdn = (EPI == LICOS_EPI_GDN || EPI == LICOS_EPI_IGDN);
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t w_bar, mma_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[256];
    __shared__ __align__(16) float beta_s[256];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int k_atoms = p.k_pad / 64, n_atoms = p.N / 64;
    const uint32_t w_bytes = (uint32_t)k_atoms * p.N * 128u;
    const uint32_t g_bytes = kGdn ? (uint32_t)n_atoms * p.N * 128u : 0u;
    uint8_t* w_s = smem;                                  // k_atoms x [N][64] bf16
    uint8_t* g_s = w_s + w_bytes;                         // n_atoms x [N][64] bf16
    uint8_t* a_s = g_s + g_bytes;                         // A tile, later the staging tile (max of the two)
    const int as_atoms = k_atoms > n_atoms ? k_atoms : n_atoms;
    float* patch = reinterpret_cast<float*>(a_s + (size_t)as_atoms * (128 * 128));

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&w_bar, 1);
        mbar_init(&mma_bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < p.N; i += kEdgeThreads) {
        bias_s[i] = p.bias ? p.bias[i] : 0.f;
        if (kGdn) beta_s[i] = p.beta[i];
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_smem, tmem_norm = tmem_base_smem + p.N;
    if (tid == 0) {
        mbar_arrive_expect_tx(&w_bar, w_bytes + g_bytes);
        for (int a = 0; a < k_atoms; ++a) tma_load_2d(w_s + (size_t)a * p.N * 128, &p.w_map, &w_bar, a * 64, 0);
        if (kGdn)
            for (int a = 0; a < n_atoms; ++a) tma_load_2d(g_s + (size_t)a * p.N * 128, &p.g_map, &w_bar, a * 64, 0);
    }

    const uint32_t idesc = umma_idesc_bf16(128, p.N);
    const uint64_t desc_hi = umma_desc_sw128(0);
    const int k_steps = (p.K + 15) / 16;       // MMAs (K = 16 each) that cover the real K
    const int n_groups8 = k_steps * 2;         // 16-byte groups of 8 k's to build per pixel
    const int patch_elems = p.C * kPatchRows * kPatchPitch;
    uint32_t mma_phase = 0;
    bool first = true;
    const int et = tid;  // epilogue threads are warps 0-3: row of the 128-row tile == TMEM lane
    const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
    const int n32 = p.N / 32;

    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int r = tile;
        const int ow0 = (r % p.tiles_w) * 16;
        r /= p.tiles_w;
        const int oh0 = (r % p.tiles_h) * 8;
        const int b = r / p.tiles_h;

        // the staging tile aliases A: the previous tile's TMA store must have read it
        if (tid == 0) tma_store_wait_read();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        // ---- input patch (fp32, zero padded) ----
        const float* xb = p.x + (size_t)b * p.C * p.H * p.W;
        for (int i = tid; i < patch_elems; i += kEdgeThreads) {
            const int col = i % kPatchPitch;
            const int rr = (i / kPatchPitch) % kPatchRows;
            const int c = i / (kPatchPitch * kPatchRows);
            const int ih = 2 * oh0 - 2 + rr, iw = 2 * ow0 - 2 + col;
            float v = 0.f;
            if (col < kPatchCols && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = __ldg(xb + ((size_t)c * p.H + ih) * p.W + iw);
            patch[i] = v;
        }
        __syncthreads();

        // ---- im2col tile A[128 pixels][k] in the 128-byte-swizzled K-major layout ----
        for (int q = tid; q < 128 * n_groups8; q += kEdgeThreads) {
            const int px = q & 127, g = q >> 7;
            const int th = px >> 4, tw = px & 15;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = g * 8 + j;
                float val = 0.f;
                if (k < p.K) {
                    const int c = k / 25, rem = k - c * 25;
                    const int kh = rem / 5, kw = rem - kh * 5;
                    val = patch[(c * kPatchRows + 2 * th + kh) * kPatchPitch + 2 * tw + kw];
                }
                v[j] = val;
            }
            *reinterpret_cast<uint4*>(a_s + (size_t)(g >> 3) * (128 * 128) + sw128_offset(px, g & 7)) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
        fence_proxy_async();
        __syncthreads();

        // ---- conv GEMM ----
        if (tid == 0) {
            if (first) mbar_wait(&w_bar, 0);
            tc_fence_after();
            const uint32_t a16 = smem_u32(a_s) >> 4, w16 = smem_u32(w_s) >> 4;
            for (int ks = 0; ks < k_steps; ++ks) {
                const uint32_t atom = ks >> 2, off = (ks & 3) * 2;
                umma_bf16(tmem_acc, desc_hi | (uint64_t)(a16 + atom * 1024 + off),
                          desc_hi | (uint64_t)(w16 + atom * ((uint32_t)p.N * 8) + off), idesc, (uint32_t)(ks > 0));
            }
            umma_commit(&mma_bar);
        }
        first = false;

        if (warp < 4) {
            mbar_wait(&mma_bar, mma_phase);
            mma_phase ^= 1u;
            tc_fence_after();
            const uint32_t t_acc = tmem_acc + lane_sel, t_norm = tmem_norm + lane_sel;
            if (kGdn) {
                for (int cc = 0; cc < n32; ++cc) {
                    float v[32];
                    tmem_ld32(t_acc + cc * 32, v);
                    tmem_ld_wait();
                    const float4* b4 = reinterpret_cast<const float4*>(bias_s + cc * 32);
                    uint32_t pk[16];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bb = b4[q];
                        const float x0 = v[4 * q] + bb.x, x1 = v[4 * q + 1] + bb.y;
                        const float x2 = v[4 * q + 2] + bb.z, x3 = v[4 * q + 3] + bb.w;
                        pk[2 * q] = pack_bf16x2(x0 * x0, x1 * x1);
                        pk[2 * q + 1] = pack_bf16x2(x2 * x2, x3 * x3);
                    }
                    uint8_t* atom = a_s + (size_t)((cc * 32) / 64) * (128 * 128);
                    const uint32_t chunk0 = ((cc * 32) % 64) / 8;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4*>(atom + sw128_offset(et, chunk0 + q)) =
                            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
                fence_proxy_async();
                tc_fence_before();
                named_bar_sync(1, 128);
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t a16 = smem_u32(a_s) >> 4, g16 = smem_u32(g_s) >> 4;
                    for (int ks = 0; ks < p.N / 16; ++ks) {
                        const uint32_t atom = ks >> 2, off = (ks & 3) * 2;
                        umma_bf16(tmem_norm, desc_hi | (uint64_t)(a16 + atom * 1024 + off),
                                  desc_hi | (uint64_t)(g16 + atom * ((uint32_t)p.N * 8) + off), idesc, (uint32_t)(ks > 0));
                    }
                    umma_commit(&mma_bar);
                }
                mbar_wait(&mma_bar, mma_phase);
                mma_phase ^= 1u;
                tc_fence_after();
            }
            for (int cc = 0; cc < n32; ++cc) {
                float v[32], nrm[32];
                tmem_ld32(t_acc + cc * 32, v);
                if (kGdn) tmem_ld32(t_norm + cc * 32, nrm);
                tmem_ld_wait();
                const float4* b4 = reinterpret_cast<const float4*>(bias_s + cc * 32);
                const float4* e4 = reinterpret_cast<const float4*>(beta_s + cc * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 bb = b4[q];
                    float x[4] = {v[4 * q] + bb.x, v[4 * q + 1] + bb.y, v[4 * q + 2] + bb.z, v[4 * q + 3] + bb.w};
                    if (kGdn) {
                        const float4 e = e4[q];
                        const float d[4] = {nrm[4 * q] + e.x, nrm[4 * q + 1] + e.y, nrm[4 * q + 2] + e.z, nrm[4 * q + 3] + e.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float rs = rsqrtf(d[i]);
                            x[i] *= (EPI == LICOS_EPI_GDN) ? rs : d[i] * rs;
                        }
                    } else if (EPI == LICOS_EPI_RELU) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) x[i] = fmaxf(x[i], 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[4 * q + i] = x[i];
                }
                uint8_t* atom = a_s + (size_t)((cc * 32) / 64) * (128 * 128);
                const uint32_t chunk0 = ((cc * 32) % 64) / 8;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(atom + sw128_offset(et, chunk0 + q)) =
                        make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
            }
            fence_proxy_async();
            named_bar_sync(1, 128);
            if (tid == 0) {
                for (int at = 0; at < n_atoms; ++at)
                    tma_store_4d(&p.out_map, a_s + (size_t)at * (128 * 128), at * 64, ow0, oh0, b);
                tma_store_commit();
            }
        }
    }
    if (tid == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base_smem, p.tmem_cols);
    }
}

}  // namespace licos
