// Weight gradient of the two 5x5 stride-2 EDGE layers of the training step (g_a[0]: image -> N, and g_s[6]: N -> image;
// reached from licos/train.py:193 `loss.backward()`), without the patch matrix ever going through HBM:
//
//   out[cs][k] += sum over (b, oh, ow) of small[b][oh][ow][cs] * image[b][c][2 oh + kh - 2][2 ow + kw - 2],  k = (c*5 + kh)*5 + kw
//
// `small` is the bf16 NHWC tensor on the many-channel side (g_a[0]: the gradient of the conv output; g_s[6]: the layer's
// input), `image` the fp32 NCHW tensor on the 1- or 3-band side (g_a[0]: the input tile; g_s[6]: the gradient of x_hat).
// The contraction runs over PIXELS, so both tcgen05 operands are MN-major (see train_kernels.cu): A = the TMA'd
// [128 px][64 ch] tiles of `small`, B = the [128 px][K] im2col tile that four builder warps assemble in shared memory from a
// TMA'd (2*8+3) x 40 x C fp32 patch -- the same construction as conv_first2.cuh's forward builders, whose K-major A tile
// has byte for byte the layout of an MN-major [pixel][64] operand.  One persistent CTA per SM accumulates its share of the
// tiles in TMEM and ends with one vector red.add flush.
//
//   warp 0      TMA: image patch ring + `small` tile ring        warp 1      MMA issuer (8 K-steps of 16 pixels per tile)
//   warps 2-5   im2col builders, then the final flush (TMEM lane == channel of `small`)
#pragma once

#include "common.cuh"

namespace licos {

constexpr int kWiPatchRows = 19, kWiPatchPitch = 40;  // fp32 patch: rows 2 oh0 - 2 .. + 18, columns 2 ow0 - 4 .. + 35
constexpr uint32_t kWiAtom = 128u * 128u;             // [128 pixels][64 columns] bf16
constexpr int kWiThreads = 192;

struct WgImageParams {
    CUtensorMap x_map;  // fp32 (W, H, C, B), box (40, 19, C, 1), no swizzle, zero fill = the conv padding
    CUtensorMap s_map;  // bf16 NHWC (Cs, OW, OH, B), box (64, 16, 8, 1), SWIZZLE_128B
    int Cs, m_blocks;   // channels of `small`; 128-channel accumulator blocks
    int k_pad;          // row pitch of `out` (licos_im2col5x5s2_kpad)
    int tiles_h, tiles_w, total_tiles;
    int slots;
    float* out;         // fp32 [Cs][k_pad], accumulated
};

template <int CIN>
struct WgImageGeom {
    static_assert(CIN == 1 || CIN == 3, "built for 1 and 3 bands");
    static constexpr int kK = CIN * 25;
    static constexpr int kAtomsB = CIN == 3 ? 2 : 1;        // K columns 0..63 | 64..79
    static constexpr int kN = CIN == 3 ? 128 : 64;          // MMA N: whole atoms; columns >= kK..: zeros / never flushed
    static constexpr int kFlush = CIN == 3 ? 80 : 32;       // columns written out (= k_pad)
    static constexpr uint32_t kPatchSlot = ((uint32_t)(CIN * kWiPatchRows * kWiPatchPitch * 4) + 127u) & ~127u;
    static constexpr uint32_t kPatchBytes = (uint32_t)(CIN * kWiPatchRows * kWiPatchPitch * 4);
};

__host__ __device__ constexpr size_t wg_image_slot_bytes(int cin, int m_blocks) {
    return (size_t)(cin == 3 ? 2 : 1) * kWiAtom + (size_t)m_blocks * 2 * kWiAtom;
}
__host__ __device__ constexpr size_t wg_image_smem_bytes(int cin, int m_blocks, int slots) {
    const size_t patch = ((size_t)cin * kWiPatchRows * kWiPatchPitch * 4 + 127) & ~(size_t)127;
    return 1024 + (size_t)slots * (wg_image_slot_bytes(cin, m_blocks) + patch);
}

template <int CIN>
__global__ void __launch_bounds__(kWiThreads, 1) wgrad_image_kernel(const __grid_constant__ WgImageParams p) {
    using G = WgImageGeom<CIN>;
    constexpr int kMaxSlots = 3;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t patch_full[kMaxSlots], patch_empty[kMaxSlots], s_full[kMaxSlots], b_full[kMaxSlots], st_empty[kMaxSlots];
    __shared__ uint64_t acc_full;
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int S = p.slots;
    const uint32_t slot_bytes = (uint32_t)wg_image_slot_bytes(CIN, p.m_blocks);
    uint8_t* patch_s = smem + (size_t)S * slot_bytes;  // slot: [B atoms][A atoms]; patches behind all slots
    const uint32_t a_off = (uint32_t)G::kAtomsB * kWiAtom;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(&patch_full[i], 1); mbar_init(&patch_empty[i], 128);
            mbar_init(&s_full[i], 1); mbar_init(&b_full[i], 128); mbar_init(&st_empty[i], 1);
        }
        mbar_init(&acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int grid = gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&p.x_map);
            tma_prefetch_desc(&p.s_map);
            const uint32_t s_bytes = (uint32_t)p.m_blocks * 2u * kWiAtom;
            int lt = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++lt) {
                const int slot = lt % S;
                const uint32_t par = (uint32_t)(lt / S) & 1u;
                int r = tile;
                const int ow0 = (r % p.tiles_w) * 16;
                r /= p.tiles_w;
                const int oh0 = (r % p.tiles_h) * 8;
                const int b = r / p.tiles_h;
                mbar_wait(&patch_empty[slot], par ^ 1u);
                mbar_arrive_expect_tx(&patch_full[slot], G::kPatchBytes);
                tma_load_4d(patch_s + (size_t)slot * G::kPatchSlot, &p.x_map, &patch_full[slot], 2 * ow0 - 4, 2 * oh0 - 2, 0, b);
                mbar_wait(&st_empty[slot], par ^ 1u);
                mbar_arrive_expect_tx(&s_full[slot], s_bytes);
                uint8_t* a0 = smem + (size_t)slot * slot_bytes + a_off;
                for (int j = 0; j < 2 * p.m_blocks; ++j)  // channel blocks past Cs are zero-filled by TMA
                    tma_load_4d(a0 + (size_t)j * kWiAtom, &p.s_map, &s_full[slot], j * 64, ow0, oh0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint64_t hi = umma_desc_mn_sw128(kWiAtom);
            const uint32_t idesc = umma_idesc_bf16(128, G::kN) | (1u << 15) | (1u << 16);  // A and B MN-major
            int lt = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++lt) {
                const int slot = lt % S;
                const uint32_t par = (uint32_t)(lt / S) & 1u;
                mbar_wait(&s_full[slot], par);
                mbar_wait(&b_full[slot], par);
                tc_fence_after();
                const uint32_t b16 = (smem_base + (uint32_t)slot * slot_bytes) >> 4, a16 = b16 + (a_off >> 4);
                for (int mb = 0; mb < p.m_blocks; ++mb) {
#pragma unroll
                    for (uint32_t r = 0; r < 8; ++r)  // K-step r = pixel row r of the 8 x 16 tile
                        umma_bf16(tmem_base + (uint32_t)mb * 128u, hi | (uint64_t)(a16 + (uint32_t)mb * ((2u * kWiAtom) >> 4) + r * 128u),
                                  hi | (uint64_t)(b16 + r * 128u), idesc, (uint32_t)(lt > 0) | r);
                }
                umma_commit(&st_empty[slot]);
            }
            umma_commit(&acc_full);
        }
    } else {
        // ===================== im2col builders (conv_first2.cuh's construction, fp32 patches) =====================
        const int u = tid - 64;  // 0..127
        int lt = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++lt) {
            const int slot = lt % S;
            const uint32_t par = (uint32_t)(lt / S) & 1u;
            uint8_t* b0 = smem + (size_t)slot * slot_bytes;
            const float* patch = reinterpret_cast<const float*>(patch_s + (size_t)slot * G::kPatchSlot);
            mbar_wait(&patch_full[slot], par);
            if constexpr (CIN == 3) {
                // a thread builds half of the K range (columns 0..39 / 40..79) of two neighbouring pixels
                const int half = u >> 6, th = (u >> 3) & 7, q = u & 7;
                const int m0 = th * 16 + 2 * q;
                const int e_base = (2 * th) * kWiPatchPitch + 4 * q + 2;  // patch columns 4q+2 .. 4q+8 of row 2 th
                uint32_t pa[20], pb[20];
                float pend_a = 0.f, pend_b = 0.f;
                auto rows = [&](auto first_row, auto n_rows) {
                    constexpr int R0 = decltype(first_row)::value, NR = decltype(n_rows)::value;
#pragma unroll
                    for (int r = R0; r < R0 + NR; ++r) {  // r = c * 5 + kh
                        const int c = r / 5, kh = r % 5;
                        const float* pr = patch + e_base + (c * kWiPatchRows + kh) * kWiPatchPitch;
                        const float2 x0 = *reinterpret_cast<const float2*>(pr);
                        const float4 x1 = *reinterpret_cast<const float4*>(pr + 2);
                        const float f[7] = {x0.x, x0.y, x1.x, x1.y, x1.z, x1.w, pr[6]};
#pragma unroll
                        for (int j = 0; j < 5; ++j) {
                            const int k = (r - R0) * 5 + j;
                            if (k & 1) {
                                pa[k >> 1] = pack_bf16x2(pend_a, f[j]);
                                pb[k >> 1] = pack_bf16x2(pend_b, f[j + 2]);
                            } else {
                                pend_a = f[j];
                                pend_b = f[j + 2];
                            }
                        }
                    }
                };
                if (half == 0) {
                    rows(std::integral_constant<int, 0>{}, std::integral_constant<int, 8>{});  // K 0..39
                } else {
                    rows(std::integral_constant<int, 8>{}, std::integral_constant<int, 7>{});  // K 40..74, then zeros
                    pa[17] = pack_bf16x2(pend_a, 0.f);
                    pb[17] = pack_bf16x2(pend_b, 0.f);
                    pa[18] = pb[18] = pa[19] = pb[19] = 0u;
                }
                mbar_arrive(&patch_empty[slot]);  // the patch is in registers now
                mbar_wait(&st_empty[slot], par ^ 1u);
#pragma unroll
                for (int gg = 0; gg < 5; ++gg) {
                    const int g = half * 5 + gg;  // 16-byte chunk of the 80-column row
                    uint8_t* base = b0 + (g < 8 ? 0u : kWiAtom);
                    const uint32_t chunk = (uint32_t)(g & 7);
                    *reinterpret_cast<uint4*>(base + sw128_offset(m0, chunk)) = make_uint4(pa[4 * gg], pa[4 * gg + 1], pa[4 * gg + 2], pa[4 * gg + 3]);
                    *reinterpret_cast<uint4*>(base + sw128_offset(m0 + 1, chunk)) = make_uint4(pb[4 * gg], pb[4 * gg + 1], pb[4 * gg + 2], pb[4 * gg + 3]);
                }
            } else {
                const int th = u >> 4, tw = u & 15;  // thread == pixel
                const int e_base = (2 * th) * kWiPatchPitch + 2 * tw + 2;
                uint32_t pa[16];
                float pend = 0.f;
#pragma unroll
                for (int kh = 0; kh < 5; ++kh) {
                    const float* pr = patch + e_base + kh * kWiPatchPitch;
                    const float2 x0 = *reinterpret_cast<const float2*>(pr), x1 = *reinterpret_cast<const float2*>(pr + 2);
                    const float f[5] = {x0.x, x0.y, x1.x, x1.y, pr[4]};
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const int k = kh * 5 + j;
                        if (k & 1) pa[k >> 1] = pack_bf16x2(pend, f[j]);
                        else pend = f[j];
                    }
                }
                pa[12] = pack_bf16x2(pend, 0.f);
                pa[13] = pa[14] = pa[15] = 0u;
                mbar_arrive(&patch_empty[slot]);
                mbar_wait(&st_empty[slot], par ^ 1u);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<uint4*>(b0 + sw128_offset(u, g)) = make_uint4(pa[4 * g], pa[4 * g + 1], pa[4 * g + 2], pa[4 * g + 3]);
            }
            fence_proxy_async();
            mbar_arrive(&b_full[slot]);
        }
        // ===================== flush: TMEM -> red.add =====================
        if (blockIdx.x < p.total_tiles) {
            const uint32_t q = (uint32_t)(warp & 3);
            const uint32_t lane_sel = (q * 32u) << 16;
            mbar_wait(&acc_full, 0);
            tc_fence_after();
            for (int mb = 0; mb < p.m_blocks; ++mb) {
                const int cs = mb * 128 + (int)q * 32 + lane;
                float* o = p.out + (size_t)cs * p.k_pad;
#pragma unroll
                for (int cc = 0; cc < G::kFlush / 16; ++cc) {
                    float v[16];
                    tmem_ld16(tmem_base + lane_sel + (uint32_t)mb * 128u + cc * 16, v);
                    tmem_ld_wait();
                    if (cs < p.Cs) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) red_add_v4(o + cc * 16 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace licos
