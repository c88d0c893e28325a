// g_s[6]: ConvTranspose2d(C -> C_out <= 4, 5x5, stride 2, padding 2, output_padding 1), bf16 NHWC in, fp32 NCHW out
// (the model's last layer).  HBM-bound: it reads the largest activation of the model.
//
// o = 2i - 2 + k  =>  output (Y, X) = (2i' + a, 2j + b) takes row taps kh = a + 2 - 2dh from input row i' + dh and
// column taps kw = b + 2 - 2dw from input column j + dw, dh, dw in {-1, 0, 1}.  The kernel splits the two directions:
//
//   columns  in the tensor core.  One GEMM per input row segment of 128 pixels, M = 128 pixels, N = 48 = (kh, b, c_out)
//            [40 used], K = 3 shifts x C: the three column shifts dw are three views of ONE TMA-loaded row of 130
//            pixels -- a SWIZZLE_128B K-major operand may start at any 128-byte row (the swizzle is a function of the
//            absolute shared-memory address; tools/desc_test.cu) -- each against its own weight slice.
//   rows     in the epilogue, for free.  The accumulators of consecutive input rows sit in a ring of 8 TMEM slots; the
//            thread that owns pixel j (TMEM lane j) sums the <= 3 row taps from the slots of rows i'-1, i', i'+1 --
//            all in its own lane -- adds the bias and stores the 2 x 2 x C_out outputs of its pixel as float2 pairs
//            (a warp store = 256 contiguous bytes of one NCHW output row).  TWO epilogue teams take alternate rows: one
//            team working the rows serially (wait, three TMEM loads, 16 outputs, 8 stores, ~2 700 cycles per row) was what
//            bounded the kernel at half its HBM roofline; a slot is handed back when its three readers have all read it.
//
// No im2col, no shared-memory staging of partial sums, no atomics; every input byte is fetched from L2 once per strip.
#pragma once

#include "common.cuh"

namespace licos {

constexpr int kN2Threads = 10 * 32;      // warp 0 loader, warp 1 MMA issuer, warps 2-5 / 6-9 the two epilogue teams
constexpr int kN2N = 48;                 // MMA N: 5 row taps x 2 column parities x 4 channel slots = 40, padded
constexpr int kN2Cpt = 4;                // channel slots
constexpr int kN2SegPx = 128;            // pixels per row segment (= MMA M)
constexpr int kN2BoxPx = kN2SegPx + 2;   // one halo pixel either side
constexpr uint32_t kN2ChunkStride = 17408;  // 130 rows x 128 B rounded up to a multiple of 1024
constexpr uint32_t kN2WTile = kN2N * 128;   // one [48][64] bf16 weight tile
constexpr int kN2AccSlots = 8, kN2AccStride = 64;
constexpr int kN2MaxSlots = 4;

struct Narrow2Params {
    CUtensorMap in_map;  // NHWC bf16 (C, W, H, B), box (64, 130, 1, 1)
    CUtensorMap w_map;   // [3 shifts x 48][C] bf16, box (64, 48)
    float* out;
    const float* bias;
    int B, H, W, C, out_c, OH, OW;
    int chunks;          // C / 64
    int strip_rows, strips, segs, total_units;
    int slots;           // A ring depth
    int relu;
    int out_mode;        // 0 = fp32, 1 = u8, 2 = u16: rint(clamp(x_hat, 0, 1) * out_scale) (LICOS_LAYOUT_NCHW_U8 / _U16)
    float out_scale;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__global__ void __launch_bounds__(kN2Threads, 1) deconv_narrow2_kernel(const __grid_constant__ Narrow2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[kN2MaxSlots], a_empty[kN2MaxSlots], acc_full[kN2AccSlots], acc_empty[kN2AccSlots], w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bias_s[kN2Cpt];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    uint8_t* w_s = smem;                                        // [shift][chunk] tiles of [48][64]
    uint8_t* a_s = w_s + (size_t)3 * p.chunks * kN2WTile;       // ring of `slots` x chunks x [130][64]
    const uint32_t slot_bytes = (uint32_t)p.chunks * kN2ChunkStride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kN2MaxSlots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kN2AccSlots; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 3 * 128); }
        mbar_init(&w_bar, 1);
        mbar_fence_init();
    }
    if (tid < kN2Cpt) bias_s[tid] = (p.bias && tid < p.out_c) ? p.bias[tid] : 0.f;
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int S = p.strip_rows;

    if (warp == 0) {
        // ===================== loader =====================
        if (lane == 0) {
            tma_prefetch_desc(&p.in_map);
            mbar_arrive_expect_tx(&w_bar, 3u * (uint32_t)p.chunks * kN2WTile);
            for (int s = 0; s < 3; ++s)
                for (int c = 0; c < p.chunks; ++c)
                    tma_load_2d(w_s + (size_t)(s * p.chunks + c) * kN2WTile, &p.w_map, &w_bar, c * 64, s * kN2N);
            uint32_t n = 0;
            for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
                int r = unit;
                const int j0 = (r % p.segs) * kN2SegPx;
                r /= p.segs;
                const int r0 = (r % p.strips) * S;
                const int b = r / p.strips;
                for (int t = 0; t < S + 2; ++t, ++n) {
                    const uint32_t slot = n % (uint32_t)p.slots;
                    mbar_wait(&a_empty[slot], ((n / (uint32_t)p.slots) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&a_full[slot], (uint32_t)p.chunks * (kN2BoxPx * 128u));
                    for (int c = 0; c < p.chunks; ++c)
                        tma_load_4d(a_s + (size_t)slot * slot_bytes + (size_t)c * kN2ChunkStride, &p.in_map, &a_full[slot],
                                    c * 64, j0 - 1, r0 - 1 + t, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        // One wait for the row's operands and one for a free accumulator slot per 24 short MMAs (N = 48: 24 cycles each)
        // left the tensor pipe idle three quarters of the time: a completed wait costs the issuing thread ~150 cycles.  The
        // NEXT row's two barriers are therefore probed (non-blocking, relaxed) before this row's MMAs are issued and read
        // after them -- the same look-ahead as conv_pair.cuh -- and the MMAs of the usual C = 128 are fully unrolled.
        mbar_wait_warp(&w_bar, 0);
        const uint32_t idesc = umma_idesc_bf16(128, kN2N);
        const uint64_t desc_hi = umma_desc_sw128(0);
        const uint32_t a16 = smem_u32(a_s) >> 4, w16 = smem_u32(w_s) >> 4;
        const uint32_t a_full0 = smem_u32(&a_full[0]), acc_empty0 = smem_u32(&acc_empty[0]);
        const uint32_t slots = (uint32_t)p.slots, slot16 = slot_bytes >> 4;
        asm volatile(".reg .pred licos_n2a, licos_n2c;");
        uint32_t slot = 0, a_par = 0, acc = 0, c_par = 1, have_a = 0, have_c = 0;
        auto wait_slow = [&](uint32_t bar, uint32_t parity) {
            if (__all_sync(0xffffffffu, mbar_try_wait_addr(bar, parity) ? 1u : 0u)) return;
            const long long t0 = clock64();
            while (!__all_sync(0xffffffffu, mbar_try_wait_addr(bar, parity) ? 1u : 0u)) {
                if (clock64() - t0 > 8000000000LL) __trap();
            }
        };
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            for (int t = 0; t < S + 2; ++t) {
                if (!__all_sync(0xffffffffu, have_a)) wait_slow(a_full0 + slot * 8u, a_par);
                if (!__all_sync(0xffffffffu, have_c)) wait_slow(acc_empty0 + acc * 8u, c_par);
                asm volatile("fence.acq_rel.cta;" ::: "memory");  // pairs with the relaxed probes
                tc_fence_after();
                const uint32_t n_slot = (slot + 1 == slots) ? 0u : slot + 1, n_apar = a_par ^ (n_slot == 0u ? 1u : 0u);
                const uint32_t n_acc = (acc + 1) & (kN2AccSlots - 1), n_cpar = c_par ^ (n_acc == 0u ? 1u : 0u);
                asm volatile("mbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 licos_n2a, [%0], %1;" ::"r"(a_full0 + n_slot * 8u),
                             "r"(n_apar)
                             : "memory");
                asm volatile("mbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 licos_n2c, [%0], %1;" ::"r"(acc_empty0 + n_acc * 8u),
                             "r"(n_cpar)
                             : "memory");
                const uint32_t d = tmem_base + acc * kN2AccStride;
                const uint32_t a_row = a16 + slot * slot16;
                if (elect_one()) {
                    if (p.chunks == 2) {
#pragma unroll
                        for (int s = 0; s < 3; ++s)  // column shift dw = s - 1: the operand starts s pixels in
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const uint32_t ab = a_row + (((uint32_t)c * kN2ChunkStride + (uint32_t)s * 128u) >> 4);
                                const uint32_t wb = w16 + (((uint32_t)(s * 2 + c) * kN2WTile) >> 4);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16(d, desc_hi | (uint64_t)(ab + 2 * k), desc_hi | (uint64_t)(wb + 2 * k), idesc,
                                              (uint32_t)((s | c | k) != 0));
                            }
                    } else {
                        uint32_t accumulate = 0;
                        for (int s = 0; s < 3; ++s)
                            for (int c = 0; c < p.chunks; ++c) {
                                const uint32_t ab = a_row + (((uint32_t)c * kN2ChunkStride + (uint32_t)s * 128u) >> 4);
                                const uint32_t wb = w16 + (((uint32_t)(s * p.chunks + c) * kN2WTile) >> 4);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    umma_bf16(d, desc_hi | (uint64_t)(ab + 2 * k), desc_hi | (uint64_t)(wb + 2 * k), idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                    }
                    umma_commit(&a_empty[slot]);
                    umma_commit(&acc_full[acc]);
                }
                __syncwarp();
                asm volatile("selp.b32 %0, 1, 0, licos_n2a;" : "=r"(have_a));
                asm volatile("selp.b32 %0, 1, 0, licos_n2c;" : "=r"(have_c));
                slot = n_slot; a_par = n_apar; acc = n_acc; c_par = n_cpar;
            }
        }
    } else {
        // ===================== epilogue: row taps + bias, straight to NCHW; two teams, alternate rows =====================
        // The GEMM of input row t is read by three output-row iterations (k = t - 2 as the upper row, t - 1 as the middle one,
        // t as the lower one), now by different teams at different times: each reader arrives on the slot's barrier (count
        // 3 x 128) and the first / last iteration of a strip arrive for the readers that do not exist at the strip's edges.
        const int team = (warp - 2) >> 2;
        const int m = (warp & 3) * 32 + lane;  // pixel of the segment == TMEM lane
        const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
        const float b0 = bias_s[0], b1 = bias_s[1], b2 = bias_s[2], b3 = bias_s[3];
        const size_t cs = (size_t)p.OH * p.OW;
        uint32_t n0 = 0;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            int r = unit;
            const int j = (r % p.segs) * kN2SegPx + m;
            r /= p.segs;
            const int r0 = (r % p.strips) * S;
            const int b = r / p.strips;
            for (int k = team; k < S; k += 2) {
                const uint32_t g_lo = n0 + k, g_mid = g_lo + 1, g_up = g_lo + 2;  // GEMMs of rows i-1, i, i+1
                mbar_wait(&acc_full[g_up % kN2AccSlots], (g_up / kN2AccSlots) & 1u);
                tc_fence_after();
                float up[16], mid[16], lo[8];
                tmem_ld16(tmem_base + lane_sel + (g_up % kN2AccSlots) * kN2AccStride, up);         // kh = 0, 1
                tmem_ld16(tmem_base + lane_sel + (g_mid % kN2AccSlots) * kN2AccStride + 16, mid);  // kh = 2, 3
                tmem_ld8(tmem_base + lane_sel + (g_lo % kN2AccSlots) * kN2AccStride + 32, lo);     // kh = 4
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_n(&acc_empty[g_lo % kN2AccSlots], k == 0 ? 3u : 1u);
                mbar_arrive_n(&acc_empty[g_mid % kN2AccSlots], 1u + (k == 0 ? 1u : 0u) + (k == S - 1 ? 1u : 0u));
                mbar_arrive_n(&acc_empty[g_up % kN2AccSlots], k == S - 1 ? 3u : 1u);
                const int i = r0 + k;
                if (i < p.H && j < p.W) {
                    // columns are [kh][b][c]: a = 0 <- kh 0 (up), 2 (mid), 4 (lo);  a = 1 <- kh 1 (up), 3 (mid)
                    float o0[8], o1[8];
                    const float bb[4] = {b0, b1, b2, b3};
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        o0[q] = up[q] + mid[q] + lo[q] + bb[q & 3];
                        o1[q] = up[8 + q] + mid[8 + q] + bb[q & 3];
                        if (p.relu) { o0[q] = fmaxf(o0[q], 0.f); o1[q] = fmaxf(o1[q], 0.f); }
                    }
                    const size_t e = ((size_t)b * p.out_c * p.OH + 2 * i) * p.OW + 2 * j;
                    if (p.out_mode == 0) {
                        float* o = p.out + e;
#pragma unroll
                        for (int c = 0; c < kN2Cpt; ++c) {
                            if (c < p.out_c) {
                                *reinterpret_cast<float2*>(o + c * cs) = make_float2(o0[c], o0[4 + c]);
                                *reinterpret_cast<float2*>(o + c * cs + p.OW) = make_float2(o1[c], o1[4 + c]);
                            }
                        }
                    } else {
                        // integer pixels: the clamp of decompress() (x_hat.clamp_(0, 1)) and the fp32 multiply + round
                        // to nearest even of a host-side `round(x_hat * int_max)`, fused into the store
                        const float sc = p.out_scale;
                        auto q = [sc](float v) { return (uint32_t)__float2int_rn(fminf(fmaxf(v, 0.f), 1.f) * sc); };
#pragma unroll
                        for (int c = 0; c < kN2Cpt; ++c) {
                            if (c < p.out_c) {
                                const uint32_t a0 = q(o0[c]), a1 = q(o0[4 + c]), b0q = q(o1[c]), b1q = q(o1[4 + c]);
                                if (p.out_mode == 1) {
                                    uint8_t* o = reinterpret_cast<uint8_t*>(p.out) + e + c * cs;
                                    *reinterpret_cast<uint16_t*>(o) = (uint16_t)(a0 | (a1 << 8));
                                    *reinterpret_cast<uint16_t*>(o + p.OW) = (uint16_t)(b0q | (b1q << 8));
                                } else {
                                    uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + e + c * cs;
                                    *reinterpret_cast<uint32_t*>(o) = a0 | (a1 << 16);
                                    *reinterpret_cast<uint32_t*>(o + p.OW) = b0q | (b1q << 16);
                                }
                            }
                        }
                    }
                }
            }
            n0 += S + 2;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// packed[s][n][cin_pad]: s = dw + 1, n = kh * 8 + b * 4 + c; kw = b + 2 - 2 dw.  w is the ConvTranspose2d weight
// (in_c, out_c, 5, 5).
__global__ void pack_weight_narrow2_kernel(const float* __restrict__ w, int out_c, int in_c, int cin_pad,
                                           __nv_bfloat16* __restrict__ packed) {
    const int64_t total = (int64_t)3 * kN2N * cin_pad;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int i = (int)(e % cin_pad);
        const int n = (int)((e / cin_pad) % kN2N);
        const int s = (int)(e / ((int64_t)cin_pad * kN2N));
        const int kh = n / 8, b = (n / 4) % 2, c = n % 4;
        const int kw = b + 2 - 2 * (s - 1);
        float v = 0.f;
        if (kh < 5 && kw >= 0 && kw < 5 && c < out_c && i < in_c) v = w[((size_t)i * out_c + c) * 25 + kh * 5 + kw];
        packed[e] = __float2bfloat16_rn(v);
    }
}

}  // namespace licos
