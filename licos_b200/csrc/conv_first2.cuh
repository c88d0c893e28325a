// g_a[0] for the reference's band counts (1 = raw split, 3 = RGB):
// Conv2d(C_in -> N, 5x5, stride 2) + bias + GDN / ReLU, fp32 NCHW in, bf16 NHWC out, as a pipelined,
// warp-specialised kernel (one persistent CTA per SM, 14 warps):
//
//   warp 0      patch loader   TMA (tiled, fp32, zero fill = the conv padding) of the (2*8+3) x 40 x C_in input
//                              patch of the next 8 x 16 output tile into a 3-slot ring
//   warp 1      MMA issuer     conv GEMM [128 px] x [K = 25 C_in + 2] x [N]: 2 (C_in = 1) or 5 (C_in = 3) tcgen05.mma;
//                              the two extra K columns carry the bias as a bf16 hi + lo pair against A = 1.0
//   warps 2-5   im2col         128 threads build the K-major, 128-byte-swizzled A tile of every output tile from 8/16-byte
//                              shared loads of the patch (fusing fp32 -> bf16 and NCHW -> NHWC into the build)
//   warps 6-9   epilogue 0     alternate tiles: TMEM -> x (packed bf16 in registers), x^2 -> smem -> gamma GEMM (in
//   warps 10-13 epilogue 1     place over the accumulator) -> x * rsqrt(beta + norm) -> bf16 -> TMA store
//
// Four TMEM accumulators (4 x N <= 512 columns), so the conv GEMM of tile t+2 never waits for the epilogue of tile t.
#pragma once

#include <type_traits>

#include "common.cuh"
#include "epilogue.cuh"

namespace licos {

// TEAMS = 2: N <= 128, 14 warps, 3 ring slots, 4 TMEM accumulators.  TEAMS = 1: N = 192 (gamma alone is 72 KB of shared
// memory), 10 warps, 2 ring slots, 2 accumulators.
__host__ __device__ constexpr int first2_threads(int teams) { return (6 + 4 * teams) * 32; }
__host__ __device__ constexpr int first2_slots(int teams) { return teams == 2 ? 3 : 2; }
__host__ __device__ constexpr int first2_bufs(int teams) { return teams == 2 ? 4 : 2; }
// patch = input rows 2*oh0-2 .. +18, columns 2*ow0-4 .. +39: a TMA box must start on a 16-byte boundary of the
// innermost dimension (measured: an unaligned start coordinate raises an illegal-instruction fault)
constexpr int kF2PatchRows = 19, kF2PatchPitch = 40;

// Input element types (N4, /root/reference/licos/raw_image_folder.py:192-196: integer tiles go straight into the first
// layer, the DN -> [0, 1] scaling happens in the im2col builders).  The patch of an integer tile starts further left so
// that the TMA box still starts on a 16-byte boundary: 8 columns for 16-bit, 16 columns for 8-bit elements.
enum { kF2InF32 = 0, kF2InU8 = 1, kF2InU16 = 2, kF2InU16Q8 = 3 };
__host__ __device__ constexpr int first2_elem_bytes(int mode) { return mode == kF2InF32 ? 4 : (mode == kF2InU8 ? 1 : 2); }
__host__ __device__ constexpr int first2_col_off(int mode) { return mode == kF2InF32 ? 4 : (mode == kF2InU8 ? 16 : 8); }
__host__ __device__ constexpr int first2_pitch(int mode) { return mode == kF2InF32 ? 40 : (mode == kF2InU8 ? 64 : 48); }

// Seven (C_in = 3) or five (C_in = 1) consecutive pixels of one patch row, starting at an EVEN element index e0 with
// e0 + 2 a multiple of four: three loads ([2][4][1] / [2][2][1] elements) whatever the element type.
template <int MODE>
struct F2Px {
    float scale;        // fl(1 / int_max) (integer modes), or fl(1 / 255) after the 8-bit re-quantisation
    uint32_t q8_magic;  // ceil(2^43 / (2 int_max)): exact (v * 510 + int_max) / (2 int_max) for v < 65536 (host-verified)
    uint32_t q8_add;    // int_max
    // Integer -> scaled float without the quarter-rate I2F: PRMT builds the float 2^23 + v (bits 0x4B000000 | v, exact for
    // v < 2^23) and ONE FFMA computes rn((2^23 + v) * scale - 2^23 * scale) = rn(v * scale): 2^23 * scale is exactly
    // representable, so this is bit-identical to float(v) * scale.
    __device__ __forceinline__ float scaled(uint32_t biased_bits) const {
        return fmaf(__uint_as_float(biased_bits), scale, -8388608.f * scale);
    }
    template <uint32_t SEL>  // PRMT selector picking the integer's byte(s) out of `word`
    __device__ __forceinline__ float pick(uint32_t word) const {
        const uint32_t biased = __byte_perm(word, 0x4B000000u, SEL);
        if (MODE == kF2InU16Q8) {  // img_as_ubyte(v / int_max) / 255: rint(v / int_max * 255) by exact integer arithmetic
            const uint32_t v = biased & 0xffffu;
            const uint64_t n = (uint64_t)(v * 510u + q8_add) * q8_magic;
            return scaled(0x4B000000u | (uint32_t)(n >> 43));
        }
        return scaled(biased);  // host-verified per int_max: bf16(this) == bf16(float32(float64(v) / int_max)) for all v
    }
    __device__ __forceinline__ void load7(const uint8_t* patch, int e0, float (&f)[7]) const {
        if (MODE == kF2InF32) {
            const float* pr = reinterpret_cast<const float*>(patch) + e0;
            const float2 a = *reinterpret_cast<const float2*>(pr);
            const float4 b = *reinterpret_cast<const float4*>(pr + 2);
            f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = b.z; f[5] = b.w; f[6] = pr[6];
        } else if (MODE == kF2InU8) {
            const uint8_t* pr = patch + e0;
            const uint32_t a = *reinterpret_cast<const uint16_t*>(pr), b = *reinterpret_cast<const uint32_t*>(pr + 2), c = pr[6];
            f[0] = pick<0x7540>(a); f[1] = pick<0x7541>(a);
            f[2] = pick<0x7540>(b); f[3] = pick<0x7541>(b); f[4] = pick<0x7542>(b); f[5] = pick<0x7543>(b);
            f[6] = pick<0x7540>(c);
        } else {
            const uint16_t* pr = reinterpret_cast<const uint16_t*>(patch) + e0;
            const uint32_t a = *reinterpret_cast<const uint32_t*>(pr), c = pr[6];
            const uint2 b = *reinterpret_cast<const uint2*>(pr + 2);
            f[0] = pick<0x7510>(a); f[1] = pick<0x7532>(a);
            f[2] = pick<0x7510>(b.x); f[3] = pick<0x7532>(b.x); f[4] = pick<0x7510>(b.y); f[5] = pick<0x7532>(b.y);
            f[6] = pick<0x7510>(c);
        }
    }
    __device__ __forceinline__ void load5(const uint8_t* patch, int e0, float (&f)[5]) const {
        if (MODE == kF2InF32) {
            const float* pr = reinterpret_cast<const float*>(patch) + e0;
            const float2 a = *reinterpret_cast<const float2*>(pr), b = *reinterpret_cast<const float2*>(pr + 2);
            f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = pr[4];
        } else if (MODE == kF2InU8) {
            const uint8_t* pr = patch + e0;
            const uint32_t a = *reinterpret_cast<const uint16_t*>(pr), b = *reinterpret_cast<const uint16_t*>(pr + 2), c = pr[4];
            f[0] = pick<0x7540>(a); f[1] = pick<0x7541>(a); f[2] = pick<0x7540>(b); f[3] = pick<0x7541>(b); f[4] = pick<0x7540>(c);
        } else {
            const uint16_t* pr = reinterpret_cast<const uint16_t*>(patch) + e0;
            const uint32_t a = *reinterpret_cast<const uint32_t*>(pr), b = *reinterpret_cast<const uint32_t*>(pr + 2), c = pr[4];
            f[0] = pick<0x7510>(a); f[1] = pick<0x7532>(a); f[2] = pick<0x7510>(b); f[3] = pick<0x7532>(b); f[4] = pick<0x7510>(c);
        }
    }
};

struct First2Params {
    CUtensorMap x_map;    // fp32 / u8 / u16 (W, H, C, B), box (pitch, 19, C, 1), no swizzle
    CUtensorMap w_map;    // [N][k_pad] bf16, box (64, N): K columns 0..63
    CUtensorMap g_map;    // gamma [N][N] bf16, box (64, N)
    CUtensorMap out_map;  // NHWC bf16 (N, OW, OH, B), box (64, 16, 8, 1)
    const __nv_bfloat16* w;  // packed weight [N][k_pad] (K columns 64.. are copied by hand)
    const float* bias;
    const float* beta;
    __nv_bfloat16* pre_out;  // optional (GDN): v = conv + bias, bf16 NHWC [B][out_h][out_w][N]
    int out_h, out_w;
    int k_pad;
    int N;
    int tiles_h, tiles_w, total_tiles;
    uint32_t tmem_cols;
    int in_mode;          // kF2In*
    float in_scale;       // see F2Px
    uint32_t q8_magic, q8_add;
};

template <int CIN>
struct First2Geom {
    static_assert(CIN == 1 || CIN == 3, "fused first layer is built for 1 and 3 bands");
    static constexpr int kK = CIN * 25;               // real K; columns kK, kK+1 carry the bias
    static constexpr int kSteps = (kK + 2 + 15) / 16;  // tcgen05.mma K steps: 2 or 5
    static constexpr int kPk = kSteps * 8;            // packed bf16 pairs per pixel row
    static constexpr bool kTail = kSteps > 4;         // K step 4 lives in the shared tail atom
    static constexpr uint32_t kPatchBytes = CIN * kF2PatchRows * kF2PatchPitch * 4;
    static constexpr uint32_t kPatchSlot = (kPatchBytes + 127u) & ~127u;
};

__host__ __device__ constexpr size_t first2_smem_bytes(int cin, int N, bool gdn, int teams) {
    const size_t n_atoms = N / 64;
    const size_t patch = ((size_t)cin * kF2PatchRows * kF2PatchPitch * 4 + 127) & ~(size_t)127;
    return 1024 + (size_t)N * 128                                  // W atom 0
           + (cin == 3 ? (size_t)(N > 128 ? N : 128) * 128 : 0)    // shared tail atom (A K-step 4 of every slot + W K-step 4)
           + (gdn ? n_atoms * N * 128 : 0)                         // gamma
           + (size_t)first2_slots(teams) * 16384                   // A ring
           + (size_t)teams * n_atoms * 16384                       // staging, one per epilogue team
           + first2_slots(teams) * patch;
}

template <int EPI, int CIN, int TEAMS>
__global__ void __launch_bounds__(first2_threads(TEAMS), 1) conv_first2_kernel(const __grid_constant__ First2Params p) {
    using G = First2Geom<CIN>;
    constexpr int kF2Threads = first2_threads(TEAMS), kF2Slots = first2_slots(TEAMS), kBufs = first2_bufs(TEAMS);
    constexpr int kXC = TEAMS == 2 ? 4 : 6;  // 32-channel chunks the epilogue is unrolled for
    constexpr bool kGdn = (EPI == LICOS_EPI_GDN || EPI == LICOS_EPI_IGDN);
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t patch_full[kF2Slots], patch_empty[kF2Slots], a_full[kF2Slots], a_empty[kF2Slots];
    __shared__ uint64_t acc_full[4], acc_empty[4], norm_full[2], w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float beta_s[256];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int N = p.N, n_atoms = N / 64;
    uint8_t* w_s = smem;
    uint8_t* tail_s = w_s + (size_t)N * 128;
    uint8_t* g_s = tail_s + (G::kTail ? (size_t)(N > 128 ? N : 128) * 128 : 0);
    uint8_t* a_s = g_s + (kGdn ? (size_t)n_atoms * N * 128 : 0);
    uint8_t* stg_s = a_s + (size_t)kF2Slots * 16384;
    uint8_t* patch_s = stg_s + (size_t)TEAMS * n_atoms * 16384;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform role dispatch
    if (tid == 0) {
        for (int i = 0; i < kF2Slots; ++i) {
            mbar_init(&patch_full[i], 1); mbar_init(&patch_empty[i], 128);
            mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
        mbar_init(&norm_full[0], 1); mbar_init(&norm_full[1], 1);
        mbar_init(&w_bar, 1);
        mbar_fence_init();
    }
    if (kGdn)
        for (int i = tid; i < N; i += kF2Threads) beta_s[i] = p.beta[i];
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // ---- resident operands: W (K columns 0..63 by TMA, the rest by hand with the bias folded in), gamma ----
    if (tid == 0) {
        mbar_arrive_expect_tx(&w_bar, (uint32_t)N * 128u + (kGdn ? (uint32_t)n_atoms * N * 128u : 0u));
        tma_load_2d(w_s, &p.w_map, &w_bar, 0, 0);
        if (kGdn)
            for (int a = 0; a < n_atoms; ++a) tma_load_2d(g_s + (size_t)a * N * 128, &p.g_map, &w_bar, a * 64, 0);
    }
    if (G::kTail) {
        if (tid < N) {  // K columns 64..79 of row n -> logical chunks 6, 7 of the tail atom
            const uint4* src = reinterpret_cast<const uint4*>(p.w + (size_t)tid * p.k_pad + 64);
            uint4 c0 = src[0], c1 = src[1];
            const float b = p.bias ? p.bias[tid] : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(b), lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            __nv_bfloat16 e[16];
            *reinterpret_cast<uint4*>(e) = c0;
            *reinterpret_cast<uint4*>(e + 8) = c1;
            e[G::kK - 64] = hi;
            e[G::kK - 64 + 1] = lo;
            *reinterpret_cast<uint4*>(tail_s + sw128_offset(tid, 6)) = *reinterpret_cast<uint4*>(e);
            *reinterpret_cast<uint4*>(tail_s + sw128_offset(tid, 7)) = *reinterpret_cast<uint4*>(e + 8);
        }
    } else {
        mbar_wait(&w_bar, 0);  // patch the bias pair into the TMA-written atom
        if (tid < N) {
            const float b = p.bias ? p.bias[tid] : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(b), lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(w_s + sw128_offset(tid, G::kK / 8));
            row[G::kK % 8] = hi;
            row[G::kK % 8 + 1] = lo;
        }
    }
    fence_proxy_async();
    __syncthreads();

    const int grid = gridDim.x;
    if (warp == 0) {
        // ===================== patch loader =====================
        if (lane == 0) {
            tma_prefetch_desc(&p.x_map);
            const int col_off = first2_col_off(p.in_mode);
            const uint32_t patch_bytes = (uint32_t)(CIN * kF2PatchRows * first2_pitch(p.in_mode) * first2_elem_bytes(p.in_mode));
            int lt = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++lt) {
                const int slot = lt % kF2Slots;
                const uint32_t par = (uint32_t)(lt / kF2Slots) & 1u;
                int r = tile;
                const int ow0 = (r % p.tiles_w) * 16;
                r /= p.tiles_w;
                const int oh0 = (r % p.tiles_h) * 8;
                const int b = r / p.tiles_h;
                mbar_wait(&patch_empty[slot], par ^ 1u);
                mbar_arrive_expect_tx(&patch_full[slot], patch_bytes);
                tma_load_4d(patch_s + (size_t)slot * G::kPatchSlot, &p.x_map, &patch_full[slot], 2 * ow0 - col_off, 2 * oh0 - 2, 0, b);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            mbar_wait(&w_bar, 0);
            const uint32_t idesc = umma_idesc_bf16(128, N);
            const uint64_t desc_hi = umma_desc_sw128(0);
            const uint32_t a16 = smem_u32(a_s) >> 4, w16 = smem_u32(w_s) >> 4, t16 = smem_u32(tail_s) >> 4;
            int lt = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++lt) {
                const int slot = lt % kF2Slots, buf = lt % kBufs;
                mbar_wait(&a_full[slot], (uint32_t)(lt / kF2Slots) & 1u);
                mbar_wait(&acc_empty[buf], ((uint32_t)(lt / kBufs) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * (uint32_t)N;
#pragma unroll
                for (int ks = 0; ks < G::kSteps; ++ks) {
                    const uint64_t ad = desc_hi | (uint64_t)(ks < 4 ? a16 + slot * 1024 + 2 * ks : t16 + 2 * slot);
                    const uint64_t bd = desc_hi | (uint64_t)(ks < 4 ? w16 + 2 * ks : t16 + 6);
                    umma_bf16(d, ad, bd, idesc, (uint32_t)(ks > 0));
                }
                umma_commit(&a_empty[slot]);
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp < 6) {
        // ===================== im2col builders =====================
        // All 128 threads work on every tile, so every thread passes through every phase of the ring barriers (a
        // waiter that skipped phases could not tell them apart by parity).  C_in = 3: a thread builds HALF of the K
        // range (K columns 0..39 or 40..79 = chunks 0..4 / 5..9) of two neighbouring pixels; C_in = 1: one pixel.
        const int u = tid - 64;  // 0..127
        auto build = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            constexpr int kPitch = first2_pitch(MODE), kShift = first2_col_off(MODE) - 4;  // element pitch, extra left margin
            const F2Px<MODE> px{p.in_scale, p.q8_magic, p.q8_add};
            int lt = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += grid, ++lt) {
                const int slot = lt % kF2Slots;
                const uint32_t par = (uint32_t)(lt / kF2Slots) & 1u;
                uint8_t* a0 = a_s + (size_t)slot * 16384;
                const uint8_t* patch = patch_s + (size_t)slot * G::kPatchSlot;
                mbar_wait(&patch_full[slot], par);
                if constexpr (CIN == 3) {
                    const int half = u >> 6, th = (u >> 3) & 7, q = u & 7;
                    const int m0 = th * 16 + 2 * q;  // rows (pixels) m0, m0 + 1 of the tile
                    const int e_base = (2 * th) * kPitch + 4 * q + 2 + kShift;  // patch columns 4q+2 .. 4q+8 of row 2 th
                    uint32_t pa[20], pb[20];
                    float pend_a = 0.f, pend_b = 0.f;
                    auto rows = [&](auto first_row, auto n_rows) {
                        constexpr int R0 = decltype(first_row)::value, NR = decltype(n_rows)::value;
#pragma unroll
                        for (int r = R0; r < R0 + NR; ++r) {  // r = c * 5 + kh: one patch row per (channel, row tap)
                            const int c = r / 5, kh = r % 5;
                            float f[7];
                            px.load7(patch, e_base + (c * kF2PatchRows + kh) * kPitch, f);
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                const int k = (r - R0) * 5 + j;  // K column relative to this half (both halves start even)
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
                        rows(std::integral_constant<int, 8>{}, std::integral_constant<int, 7>{});  // K 40..74
                        // K 75, 76: 1.0 against the bias hi / lo rows of W, then zeros
                        pa[17] = pack_bf16x2(pend_a, 1.f);
                        pb[17] = pack_bf16x2(pend_b, 1.f);
                        pa[18] = pb[18] = pack_bf16x2(1.f, 0.f);
                        pa[19] = pb[19] = 0u;
                    }
                    mbar_arrive(&patch_empty[slot]);  // the patch is in registers now
                    mbar_wait(&a_empty[slot], par ^ 1u);
#pragma unroll
                    for (int gg = 0; gg < 5; ++gg) {
                        const int g = half * 5 + gg;  // 16-byte chunk of the 80-column row
                        uint8_t* base = (g < 8) ? a0 : tail_s;
                        const uint32_t chunk = (g < 8) ? (uint32_t)g : (uint32_t)(2 * slot + (g - 8));
                        *reinterpret_cast<uint4*>(base + sw128_offset(m0, chunk)) =
                            make_uint4(pa[4 * gg], pa[4 * gg + 1], pa[4 * gg + 2], pa[4 * gg + 3]);
                        *reinterpret_cast<uint4*>(base + sw128_offset(m0 + 1, chunk)) =
                            make_uint4(pb[4 * gg], pb[4 * gg + 1], pb[4 * gg + 2], pb[4 * gg + 3]);
                    }
                } else {
                    const int th = u >> 4, tw = u & 15;  // thread == pixel (row u of the tile)
                    const int e_base = (2 * th) * kPitch + 2 * tw + 2 + kShift;  // patch columns 2tw+2 .. 2tw+6 of row 2 th
                    uint32_t pa[16];
                    float pend = 0.f;
#pragma unroll
                    for (int kh = 0; kh < 5; ++kh) {
                        float f[5];
                        px.load5(patch, e_base + kh * kPitch, f);
#pragma unroll
                        for (int j = 0; j < 5; ++j) {
                            const int k = kh * 5 + j;
                            if (k & 1) pa[k >> 1] = pack_bf16x2(pend, f[j]);
                            else pend = f[j];
                        }
                    }
                    pa[12] = pack_bf16x2(pend, 1.f);  // K 24, then 1.0 (bias hi)
                    pa[13] = pack_bf16x2(1.f, 0.f);   // 1.0 (bias lo), zeros
                    pa[14] = pa[15] = 0u;
                    mbar_arrive(&patch_empty[slot]);
                    mbar_wait(&a_empty[slot], par ^ 1u);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        *reinterpret_cast<uint4*>(a0 + sw128_offset(u, g)) = make_uint4(pa[4 * g], pa[4 * g + 1], pa[4 * g + 2], pa[4 * g + 3]);
                }
                fence_proxy_async();
                mbar_arrive(&a_full[slot]);
            }
        };
        switch (p.in_mode) {  // warp-uniform: one branch per launch, the tile loop lives inside
            case kF2InU8: build(std::integral_constant<int, kF2InU8>{}); break;
            case kF2InU16: build(std::integral_constant<int, kF2InU16>{}); break;
            case kF2InU16Q8: build(std::integral_constant<int, kF2InU16Q8>{}); break;
            default: build(std::integral_constant<int, kF2InF32>{}); break;
        }
    } else {
        // ===================== epilogue teams =====================
        const int team = (warp - 6) >> 2;
        const int row = (warp & 3) * 32 + lane;  // TMEM lane == row of the tile; warp % 4 selects the lane quadrant
        const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
        const bool leader = (warp & 3) == 0 && lane == 0;
        uint8_t* stg = stg_s + (size_t)team * n_atoms * 16384;
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint32_t stg16 = smem_u32(stg) >> 4, g16 = smem_u32(g_s) >> 4;
        const int n32 = N / 32;
        uint32_t nit = 0;
        for (int lt = team;; lt += TEAMS) {
            const int tile = blockIdx.x + lt * grid;
            if (tile >= p.total_tiles) break;
            int r = tile;
            const int ow0 = (r % p.tiles_w) * 16;
            r /= p.tiles_w;
            const int oh0 = (r % p.tiles_h) * 8;
            const int b = r / p.tiles_h;
            const int buf = lt % kBufs;
            const uint32_t t_acc = tmem_base + lane_sel + (uint32_t)buf * (uint32_t)N;

            // (the wait for this team's previous TMA store to have read the staging tile comes after the first accumulator piece
            // has been loaded and worked on, right before the first write to the tile: it overlaps that work)
            auto staging_free = [&]() {
                if (leader) tma_store_wait_read();
                named_bar_sync(1 + team, 128);
            };
            mbar_wait(&acc_full[buf], (uint32_t)(lt / kBufs) & 1u);
            tc_fence_after();

            uint32_t xs[kGdn ? kXC * 16 : 1];
            if (kGdn) {
#pragma unroll
                for (int cc = 0; cc < kXC; ++cc) {
                    if (cc < n32) {
                        float v[32];
                        tmem_ld32(t_acc + cc * 32, v);
                        tmem_ld_wait();
                        uint32_t sq[16];
                        gdn_stage1_32<false>(v, nullptr, xs + cc * 16, sq);
                        if (cc == 0) staging_free();
                        store_row32(stg, row, cc, sq);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                named_bar_sync(1 + team, 128);
                if ((warp & 3) == 0) {  // the team's first warp, converged after the barrier
                    tc_fence_after();
                    if (elect_one()) {
                        issue_gamma_gemm_n(tmem_base + (uint32_t)buf * (uint32_t)N, stg16, g16, (uint32_t)N, idesc);
                        umma_commit(&norm_full[team]);
                    }
                    __syncwarp();
                }
                mbar_wait(&norm_full[team], nit & 1u);
                tc_fence_after();
                ++nit;
            }
            const int oh = oh0 + (row >> 4), ow = ow0 + (row & 15);  // staging rows run w-fastest over the 8 x 16 tile
            __nv_bfloat16* pre_px = (kGdn && p.pre_out && oh < p.out_h && ow < p.out_w)
                                        ? p.pre_out + (((size_t)b * p.out_h + oh) * p.out_w + ow) * N
                                        : nullptr;
#pragma unroll
            for (int cc = 0; cc < kXC; ++cc) {
                if (cc < n32) {
                    float v[32];
                    tmem_ld32(t_acc + cc * 32, v);
                    tmem_ld_wait();
                    uint32_t out[16];
                    if (kGdn) {
                        if (pre_px) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                reinterpret_cast<uint4*>(pre_px + cc * 32)[q] =
                                    make_uint4(xs[cc * 16 + 4 * q], xs[cc * 16 + 4 * q + 1], xs[cc * 16 + 4 * q + 2], xs[cc * 16 + 4 * q + 3]);
                        }
                        gdn_stage2_32<EPI == LICOS_EPI_IGDN>(v, beta_s + cc * 32, xs + cc * 16, out);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float x0 = v[2 * i], x1 = v[2 * i + 1];
                            if (EPI == LICOS_EPI_RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                            out[i] = pack_bf16x2(x0, x1);
                        }
                    }
                    if (!kGdn && cc == 0) staging_free();
                    store_row32(stg, row, cc, out);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
            fence_proxy_async();
            named_bar_sync(1 + team, 128);
            if (leader) {
                for (int at = 0; at < n_atoms; ++at)
                    tma_store_4d(&p.out_map, stg + (size_t)at * 16384, at * 64, ow0, oh0, b);
                tma_store_commit();
            }
        }
        if (leader) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

}  // namespace licos
