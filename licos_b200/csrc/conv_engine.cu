// Implicit-GEMM convolution engine for sm_100a: strided 5x5 conv, 5x5 stride-2 transposed conv (as four
// output-parity sub-convolutions), 3x3 stride-1 conv, each with its epilogue (bias, GDN / IGDN, ReLU)
// fused.  One persistent, warp-specialised kernel, one CTA per SM, 12 warps:
//
//   warp 0     A producer   TMA loads "slabs" of the NHWC bf16 activation: for one column tap kw and one
//                           64-channel chunk, a box of (TH+2) rows x 16 columns x 64 channels.  Every row
//                           tap kh that shares this kw reads the SAME slab at a row offset (a multiple of
//                           2 KB, so the 128-byte swizzle phase is unchanged) -> each activation byte is
//                           fetched 5x (not 25x) per 5x5 layer.
//   warp 1     B producer   TMA loads one [N x 64] weight tile per (tap, chunk) (warp-uniform loop, tap list in
//                           shared memory); loads gamma once, resident for the whole kernel.
//   warp 2     MMA issuer   warp-uniform loop; the elected lane issues the tcgen05.mma (M=128, N<=256, K=16) of a
//                           tap back to back into two TMEM accumulators (two accumulator SETS: the epilogue of one
//                           pass overlaps the mainloop of the next).
//   warps 4-11 epilogue     two teams of 4 warps taking alternate accumulators: TMEM -> registers -> (+bias,
//                           square) -> smem -> gamma GEMM in place over the accumulator -> rsqrt / sqrt -> bf16
//                           NHWC via TMA store, or fp32 NCHW direct stores.
//
// Stride-2 addressing is done with four parity views of the input (conv) or output (deconv) tensor,
// each an ordinary tiled TMA descriptor; padding and ragged edges are TMA out-of-bounds zero fill /
// store clipping.  The first layer (fp32 NCHW in, K = 25 C_in) and the last layer (<= 4 output channels) have
// their own kernels (conv_first2.cuh, deconv_narrow2.cuh; conv_edge.cuh is the generic first-layer fallback).
// Replaces SURVEY.md section 8a rows A3, A4, A5.  DESIGN.md section 4.0 lists the measurements behind the structure.
#include "common.cuh"
#include "conv_edge.cuh"
#include "epilogue.cuh"
#include "conv_first2.cuh"
#include "deconv_narrow2.cuh"

#include <math.h>
#include <stdlib.h>
#include <mutex>
#include <string.h>

namespace licos {

constexpr int kMaxSlabs = 10;
constexpr int kMaxTaps = 10;
constexpr int kMaxPasses = 4;
constexpr int kTileW = 16;    // tile columns
constexpr int kAccRows = 8;   // tile rows per accumulator: 8 x 16 = 128 = MMA M
constexpr int kKChunk = 64;   // channels per K chunk = one 128-byte swizzle atom of bf16
constexpr uint32_t kRowBytes = kTileW * kKChunk * 2;  // one slab row: 2 KB
constexpr int kThreads = 384;  // A producer, B producer, MMA issuer, (idle), 2 x 4 epilogue warps
constexpr int kMaxSA = 4, kMaxSB = 8;
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxDynSmem = 228000;  // 227 KB opt-in limit minus ~3.4 KB of static shared memory

struct Tap {
    int8_t row_off;  // slab row of the tile's first row for this tap
    int8_t group;    // accumulator group (output parity for the merged narrow deconv)
    int16_t w_tap;   // tap index into the packed weight
};
struct Slab {
    int8_t in_map;
    int8_t dw;  // column shift of the slab relative to the tile
    int8_t n_taps;
    int8_t pad_;
    Tap taps[kMaxTaps];
};
struct Pass {
    int8_t n_slabs;
    int8_t n_groups;
    int8_t pad_[2];
    int8_t out_map[4];  // per group: TMA store map (NHWC output)
    int8_t dy[4];       // per group: output position = grid position * out_s + (dy, dx)
    int8_t dx[4];
    Slab slabs[kMaxSlabs];
};

struct ConvParams {
    CUtensorMap in_maps[4];
    CUtensorMap out_maps[4];
    CUtensorMap w_map;
    CUtensorMap g_map;
    Pass passes[kMaxPasses];
    int n_passes;
    int batch;
    int grid_h, grid_w;  // extent of the position grid the M tile walks over
    int tiles_h, tiles_w;
    int n_acc;       // 128-row sub-tiles per CTA tile (tile rows = 8 * n_acc)
    int N;           // MMA N
    int n_split;     // output-channel splits (out_c > 256)
    int cin_chunks;  // padded input channels / 64
    int w_rows_per_tap;
    int epilogue;
    int out_layout;
    int out_c, out_h, out_w, out_s;
    float* out_f32;
    __nv_bfloat16* pre_out;  // optional (GDN epilogues): v = conv + bias, bf16 NHWC -- the backward pass's saved input
    const float* bias;
    const float* beta;
    int sa, sb;
    uint32_t a_slot_bytes, b_slot_bytes, staging_bytes, gamma_bytes;
    unsigned long long pass_info[kMaxPasses];  // lean issue loop: per slab 5 bits {n_taps-1 : 2, first row_off : 2, descending : 1}
    // CTA-pair kernel (conv_pair.cuh): the geometry of a tile and the taps of every slab.  wide = 1: an accumulator is a
    // block of 16 rows x 8 columns and a slab is one (16 + 2) x (8 n_acc + 2)-pixel window of a parity view that serves
    // EVERY tap of that view through row- and column-shifted operand starts (the stride between 8-row groups is the slab's
    // row pitch); wide = 0: the 8-row x 16-column blocks and 16-pixel-wide, per-column-tap slabs of conv_igemm_kernel.
    unsigned long long tap_list[kMaxPasses][kMaxSlabs];  // [0:4) taps of the slab, then (row_off : 2, col_off : 2) per tap
    int wide, tile_h, tile_w;
    uint32_t a_tx_bytes;                      // bytes one slab load delivers (the ring slot is that rounded up to 1 KB)
    uint32_t a_pitch16, a_sbo16, acc_step16;  // slab row pitch, 8-row-group stride, operand offset of accumulator 1 (16-byte units)
    int lean;           // 1 = every slab's taps walk consecutive slab rows (pass_info valid), 0 = generic per-tap table walk
    int n_teams;        // epilogue teams that take jobs (2, or 1 when two staging tiles do not fit in shared memory)
    int jobs_per_pass;  // accumulators per pass (n_groups * n_acc): each is one epilogue job
    int n_buf;  // TMEM accumulator sets (2 = the epilogue of pass p overlaps the mainloop of pass p+1)
    int total_tiles;
    unsigned long long* dbg;  // optional per-CTA cycle probes (licos_internal_set_conv_probe, development only)
    int dbg_flags;            // LICOS_DBG_FLAGS experiments: 1 = A slabs loaded only once per ring slot, 2 = same for weights
};

// probe slots (per CTA, 16 x u64)
enum { DBG_PA_WAIT = 0, DBG_PB_WAIT, DBG_MMA_A, DBG_MMA_B, DBG_MMA_ACC, DBG_MMA_X2, DBG_MMA_TOTAL, DBG_EPI_ACC,
       DBG_EPI_S1, DBG_EPI_NORM, DBG_EPI_S2, DBG_EPI_STORE, DBG_EPI_TOTAL, DBG_TILES, DBG_PA_TOTAL, DBG_PB_TOTAL };
#define PROBE_T0() const long long _t0 = p.dbg ? clock64() : 0
#define PROBE_ADD(var) do { if (p.dbg) (var) += clock64() - _t0; } while (0)

struct TileCoord {
    int b, gh0, gw0, ns;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
    TileCoord t;
    t.ns = tile % p.n_split;
    int r = tile / p.n_split;
    t.gw0 = (r % p.tiles_w) * kTileW;
    r /= p.tiles_w;
    t.gh0 = (r % p.tiles_h) * (kAccRows * p.n_acc);
    t.b = r / p.tiles_h;
    return t;
}

// ring cursor: slot index + phase bit, advanced without integer division
struct Ring {
    uint32_t slot = 0, phase = 0;
    __device__ __forceinline__ void advance(uint32_t n) {
        if (++slot == n) { slot = 0; phase ^= 1u; }
    }
};

// XC = number of 32-channel chunks the epilogue is unrolled for (4: N <= 128, 6: N <= 192, 8: N <= 256)
template <int EPI, bool OUT_NHWC, int XC>
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
    constexpr bool kGdn = (EPI == LICOS_EPI_GDN || EPI == LICOS_EPI_IGDN);
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], b_full[kMaxSB], b_empty[kMaxSB];
    __shared__ uint64_t acc_full[2], acc_empty[2], norm_full[2], g_full;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[512];
    __shared__ __align__(16) float beta_s[256];
    __shared__ int16_t w_taps_s[kMaxPasses][kMaxSlabs * kMaxTaps];  // weight tap of the i-th B tile of a pass (per chunk)
    __shared__ int n_taps_s[kMaxPasses];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + (size_t)p.sa * p.a_slot_bytes;
    uint8_t* staging_all = b_ring + (size_t)p.sb * p.b_slot_bytes;  // one tile per epilogue team
    uint8_t* gamma_s = staging_all + (size_t)p.n_teams * p.staging_bytes;   // resident gamma: (N / 64) atoms of [N][64] bf16

    // the warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps the
    // issue loops' state in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < p.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128u * (uint32_t)p.jobs_per_pass);
            mbar_init(&norm_full[i], 1);
        }
        mbar_init(&g_full, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < p.N * p.n_split; i += kThreads) bias_s[i] = (p.bias && i < p.out_c) ? p.bias[i] : 0.f;
    if (kGdn)
        for (int i = threadIdx.x; i < p.N; i += kThreads) beta_s[i] = p.beta[i];
    if (threadIdx.x < p.n_passes) {
        const Pass& ps = p.passes[threadIdx.x];
        int n = 0;
        for (int sl = 0; sl < ps.n_slabs; ++sl)
            for (int k = 0; k < ps.slabs[sl].n_taps; ++k) w_taps_s[threadIdx.x][n++] = ps.slabs[sl].taps[k].w_tap;
        n_taps_s[threadIdx.x] = n;
    }
    if (warp == 2) {
        tmem_alloc(&tmem_base_smem, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0 && lane == 0) {
        // ===================== A producer =====================
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.in_maps[i]);
        Ring ra;
        int n_loaded = 0;
        long long w_wait = 0;
        const long long t_begin = p.dbg ? clock64() : 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            for (int pi = 0; pi < p.n_passes; ++pi) {
                const Pass& ps = p.passes[pi];
                for (int c = 0; c < p.cin_chunks; ++c) {
                    for (int s = 0; s < ps.n_slabs; ++s) {
                        const Slab& sl = ps.slabs[s];
                        { PROBE_T0(); mbar_wait(&a_empty[ra.slot], ra.phase ^ 1u); PROBE_ADD(w_wait); }
                        if ((p.dbg_flags & 1) && n_loaded >= p.sa) { mbar_arrive(&a_full[ra.slot]); ra.advance(p.sa); continue; }
                        ++n_loaded;
                        mbar_arrive_expect_tx(&a_full[ra.slot], p.a_slot_bytes);
                        tma_load_4d(a_ring + (size_t)ra.slot * p.a_slot_bytes, &p.in_maps[sl.in_map], &a_full[ra.slot],
                                    c * kKChunk, t.gw0 + sl.dw, t.gh0 - 1, t.b);
                        ra.advance(p.sa);
                    }
                }
            }
        }
        if (p.dbg) {
            p.dbg[blockIdx.x * 16 + DBG_PA_WAIT] = w_wait;
            p.dbg[blockIdx.x * 16 + DBG_PA_TOTAL] = clock64() - t_begin;
        }
    } else if (warp == 1) {
        // ===================== B producer (warp-uniform loop, one elected lane issues the TMA) =====================
        if (lane == 0) {
            tma_prefetch_desc(&p.w_map);
            if (kGdn) {  // gamma stays resident for the whole kernel
                tma_prefetch_desc(&p.g_map);
                mbar_arrive_expect_tx(&g_full, p.gamma_bytes);
                for (int gc = 0; gc < p.N / kKChunk; ++gc)
                    tma_load_2d(gamma_s + (size_t)gc * p.N * 128, &p.g_map, &g_full, gc * kKChunk, 0);
            }
        }
        __syncwarp();
        uint32_t slot = 0, phase = 0;
        int n_loaded = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int ns_row = (tile % p.n_split) * p.N;
            for (int pi = 0; pi < p.n_passes; ++pi) {
                const int nt = n_taps_s[pi];
                for (int c = 0; c < p.cin_chunks; ++c) {
                    for (int i = 0; i < nt; ++i) {
                        const int row = (int)w_taps_s[pi][i] * p.w_rows_per_tap + ns_row;
                        mbar_wait_warp(&b_empty[slot], phase ^ 1u);
                        if (elect_one()) {
                            if ((p.dbg_flags & 2) && n_loaded >= p.sb) {
                                mbar_arrive(&b_full[slot]);
                            } else {
                                mbar_arrive_expect_tx(&b_full[slot], p.b_slot_bytes);
                                tma_load_2d(b_ring + (size_t)slot * p.b_slot_bytes, &p.w_map, &b_full[slot], c * kKChunk, row);
                            }
                        }
                        __syncwarp();
                        ++n_loaded;
                        slot = (slot + 1 == (uint32_t)p.sb) ? 0u : slot + 1;
                        phase ^= (slot == 0u) ? 1u : 0u;
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== MMA issuer =====================
        // Measured (tools/mma_bench3-6.cu): tcgen05.mma issue is synchronous with the issuing thread's instruction
        // stream.  The pipe sustains its 64-cycle rate (M128 N128 K16) only while the next MMA is presented within a few
        // cycles; every instruction of the issuing thread between two MMAs -- ring bookkeeping, descriptor arithmetic,
        // ~90 cycles per (completed) mbarrier wait -- runs at full latency (one warp, nothing to hide it behind) and
        // idles the pipe, while a commit is free.  So: a WARP-UNIFORM loop with the MMAs of a tap issued back to back by the
        // elected lane (no R2UR / ELECT sequence per MMA), alternating between the two accumulators; descriptors advance
        // by register adds (the taps of a slab walk consecutive slab rows: pass_info); no table loads, no probes.
        // Variants that were measured and dropped: two issuers, one per accumulator (fall into lock-step on the shared
        // barriers); 2 / 4 issuers taking alternate taps with a shared turn counter (the hand-over costs as much as the
        // bookkeeping it hides); one test_wait probing all ring slots at once (the waits were not the bottleneck).
        const uint32_t idesc = umma_idesc_bf16(128, p.N);
        const uint64_t desc_hi = umma_desc_sw128(0);
        const uint32_t a_ring_addr = smem_u32(a_ring) >> 4, b_ring_addr = smem_u32(b_ring) >> 4;
        const uint32_t a_slot16 = p.a_slot_bytes >> 4, b_slot16 = p.b_slot_bytes >> 4;
        const uint32_t n_acc = p.n_acc, N = p.N, sa = p.sa, sb = p.sb;
        constexpr uint32_t kAccStep16 = (kAccRows * kRowBytes) >> 4;
        uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0, pit = 0;
        const long long t_begin = p.dbg ? clock64() : 0;
        long long n_tiles = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            ++n_tiles;
            for (int pi = 0; pi < p.n_passes; ++pi, ++pit) {
                const Pass& ps = p.passes[pi];
                const uint32_t buf = pit % (uint32_t)p.n_buf;
                mbar_wait_warp(&acc_empty[buf], ((pit / (uint32_t)p.n_buf) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_set = tmem_base + buf * (uint32_t)(ps.n_groups * (int)n_acc) * N;
                if (p.lean) {
                    const int n_slabs = ps.n_slabs;
                    uint32_t accumulate = 0;
                    for (int c = 0; c < p.cin_chunks; ++c) {
                        unsigned long long info = p.pass_info[pi];
                        for (int s = 0; s < n_slabs; ++s) {
                            const int nt = (int)(info & 3u) + 1;
                            uint32_t a_lo = a_ring_addr + a_slot * a_slot16 + ((uint32_t)(info >> 2) & 3u) * (kRowBytes >> 4);
                            const uint32_t a_step = (info & 16u) ? 0u - (kRowBytes >> 4) : (kRowBytes >> 4);
                            info >>= 5;
                            mbar_wait_warp(&a_full[a_slot], a_phase);
                            for (int k = 0; k < nt; ++k) {
                                mbar_wait_warp(&b_full[b_slot], b_phase);
                                tc_fence_after();
                                const uint64_t bd = desc_hi | (uint64_t)(b_ring_addr + b_slot * b_slot16);
                                const uint64_t ad = desc_hi | (uint64_t)a_lo;
                                if (elect_one()) {
                                    if (n_acc == 2) {
                                        umma_bf16(tmem_set, ad, bd, idesc, accumulate);
                                        umma_bf16(tmem_set + N, ad + kAccStep16, bd, idesc, accumulate);
#pragma unroll
                                        for (uint32_t ks = 1; ks < 4; ++ks) {
                                            umma_bf16(tmem_set, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                                            umma_bf16(tmem_set + N, ad + kAccStep16 + 2 * ks, bd + 2 * ks, idesc, 1u);
                                        }
                                    } else {
                                        umma_bf16(tmem_set, ad, bd, idesc, accumulate);
#pragma unroll
                                        for (uint32_t ks = 1; ks < 4; ++ks) umma_bf16(tmem_set, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                                    }
                                    umma_commit(&b_empty[b_slot]);
                                }
                                __syncwarp();
                                accumulate = 1;
                                a_lo += a_step;
                                b_slot = (b_slot + 1 == sb) ? 0u : b_slot + 1;
                                b_phase ^= (b_slot == 0u) ? 1u : 0u;
                            }
                            if (elect_one()) umma_commit(&a_empty[a_slot]);
                            __syncwarp();
                            a_slot = (a_slot + 1 == sa) ? 0u : a_slot + 1;
                            a_phase ^= (a_slot == 0u) ? 1u : 0u;
                        }
                    }
                } else {
                    // generic walk of the tap table (merged multi-group passes): rare shapes, not tuned
                    uint32_t touched = 0;
                    for (int c = 0; c < p.cin_chunks; ++c) {
                        for (int s = 0; s < ps.n_slabs; ++s) {
                            const Slab& sl = ps.slabs[s];
                            mbar_wait_warp(&a_full[a_slot], a_phase);
                            const uint32_t a_slab = a_ring_addr + a_slot * a_slot16;
                            const int n_taps = sl.n_taps;
                            for (int k = 0; k < n_taps; ++k) {
                                const Tap tp = sl.taps[k];
                                mbar_wait_warp(&b_full[b_slot], b_phase);
                                tc_fence_after();
                                const uint64_t bd = desc_hi | (uint64_t)(b_ring_addr + b_slot * b_slot16);
                                uint64_t ad = desc_hi | (uint64_t)(a_slab + (uint32_t)tp.row_off * (kRowBytes >> 4));
                                uint32_t acc = (uint32_t)tp.group * n_acc;
                                if (elect_one()) {
                                    for (uint32_t a = 0; a < n_acc; ++a, ++acc, ad += kAccStep16) {
                                        const uint32_t d = tmem_set + acc * N;
                                        umma_bf16(d, ad, bd, idesc, (touched >> acc) & 1u);
                                        umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                                        umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
                                        umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
                                    }
                                    umma_commit(&b_empty[b_slot]);
                                }
                                __syncwarp();
                                for (uint32_t a = 0; a < n_acc; ++a) touched |= 1u << ((uint32_t)tp.group * n_acc + a);
                                b_slot = (b_slot + 1 == sb) ? 0u : b_slot + 1;
                                b_phase ^= (b_slot == 0u) ? 1u : 0u;
                            }
                            if (elect_one()) umma_commit(&a_empty[a_slot]);
                            __syncwarp();
                            a_slot = (a_slot + 1 == sa) ? 0u : a_slot + 1;
                            a_phase ^= (a_slot == 0u) ? 1u : 0u;
                        }
                    }
                }
                if (elect_one()) umma_commit(&acc_full[buf]);
                __syncwarp();
            }
        }
        if (p.dbg && lane == 0) {
            unsigned long long* d = p.dbg + blockIdx.x * 16;
            d[DBG_MMA_A] = 0; d[DBG_MMA_B] = 0; d[DBG_MMA_ACC] = 0; d[DBG_MMA_X2] = 0;
            d[DBG_MMA_TOTAL] = clock64() - t_begin; d[DBG_TILES] = n_tiles;
        }
    } else if (warp >= 4) {
        // ===================== epilogue: two teams of 4 warps, alternate accumulators =====================
        const int team = (warp - 4) >> 2;
        const int et = (warp & 3) * 32 + lane;  // == TMEM lane == row of the 128-row sub-tile
        const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
        const bool leader = et == 0;
        const int th = et / kTileW, tw = et % kTileW;
        uint8_t* staging = staging_all + (size_t)team * p.staging_bytes;
        uint32_t pit = 0, nit = 0, job = 0;
        const int n32 = p.N / 32;
        const uint32_t idesc = umma_idesc_bf16(128, p.N);
        const uint32_t staging16 = smem_u32(staging) >> 4, gamma16 = smem_u32(gamma_s) >> 4;
        bool gamma_ready = false;
        const bool probe = p.dbg && leader && team == 0;
        long long e_acc = 0, e_s1 = 0, e_norm = 0, e_s2 = 0, e_store = 0;
        const long long t_begin = probe ? clock64() : 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            const float* bias_t = bias_s + t.ns * p.N;
            for (int pi = 0; pi < p.n_passes; ++pi, ++pit) {
                const Pass& ps = p.passes[pi];
                const uint32_t buf = pit % (uint32_t)p.n_buf;
                const uint32_t tmem_set = tmem_base + lane_sel + buf * (uint32_t)(ps.n_groups * p.n_acc) * p.N;
                bool waited = false;
                for (int g = 0; g < ps.n_groups; ++g) {
                    for (int a = 0; a < p.n_acc; ++a) {
                        if (((job++) & (uint32_t)(p.n_teams - 1)) != (uint32_t)team) continue;
                        if (!waited) {
                            const long long _t = probe ? clock64() : 0;
                            mbar_wait(&acc_full[buf], (pit / (uint32_t)p.n_buf) & 1u);
                            if (probe) e_acc += clock64() - _t;
                            tc_fence_after();
                            waited = true;
                        }
                        const uint32_t acc = (uint32_t)g * p.n_acc + a;
                        const uint32_t t_acc = tmem_set + acc * p.N;

                        if (OUT_NHWC || kGdn) {
                            // this team's previous TMA store must have finished reading `staging` before it is rewritten
                            const long long _st = probe ? clock64() : 0;
                            if (OUT_NHWC && leader) tma_store_wait_read();
                            named_bar_sync(1 + team, 128);
                            if (probe) e_store += clock64() - _st;
                        }
                        uint32_t xs[kGdn ? XC * 16 : 1];  // v = acc + bias kept as packed bf16 pairs
                        if (kGdn) {
                            // stage 1: v^2 (bf16) -> staging = A operand of the gamma GEMM; the GEMM then overwrites
                            // the accumulator IN PLACE with the norm (no second TMEM region -> room to double-buffer)
                            const long long _s1 = probe ? clock64() : 0;
#pragma unroll
                            for (int cc = 0; cc < XC; ++cc) {
                                if (cc < n32) {
                                    float v[32];
                                    tmem_ld32(t_acc + cc * 32, v);
                                    tmem_ld_wait();
                                    uint32_t sq[16];
                                    gdn_stage1_32<true>(v, bias_t + cc * 32, xs + cc * 16, sq);
                                    store_row32(staging, et, cc, sq);
                                }
                            }
                            fence_proxy_async();
                            tc_fence_before();
                            named_bar_sync(1 + team, 128);
                            if ((warp & 3) == 0) {  // the team's first warp, converged after the barrier
                                if (!gamma_ready) { mbar_wait(&g_full, 0); gamma_ready = true; }
                                tc_fence_after();
                                if (elect_one()) {
                                    issue_gamma_gemm_n(tmem_set - lane_sel + acc * p.N, staging16, gamma16, (uint32_t)p.N, idesc);
                                    umma_commit(&norm_full[team]);
                                }
                                __syncwarp();
                            }
                            if (probe) e_s1 += clock64() - _s1;
                            const long long _n = probe ? clock64() : 0;
                            mbar_wait(&norm_full[team], nit & 1u);
                            if (probe) e_norm += clock64() - _n;
                            tc_fence_after();
                            ++nit;
                        }

                        // stage 2: activation, then write out
                        const long long _s2 = probe ? clock64() : 0;
                        const int gh = t.gh0 + a * kAccRows + th, gw = t.gw0 + tw;
                        const int oh = gh * p.out_s + ps.dy[g], ow = gw * p.out_s + ps.dx[g];
                        const bool in_range = gh < p.grid_h && gw < p.grid_w && oh < p.out_h && ow < p.out_w;
                        const size_t cs = (size_t)p.out_h * p.out_w;
                        float* o = OUT_NHWC ? nullptr
                                            : p.out_f32 + ((size_t)t.b * p.out_c * p.out_h + oh) * p.out_w + ow +
                                                  (size_t)(t.ns * p.N) * cs;
                        const int c_left = p.out_c - t.ns * p.N;  // valid channels from this split's base
                        __nv_bfloat16* pre_px = (kGdn && p.pre_out && in_range)
                                                    ? p.pre_out + (((size_t)t.b * p.out_h + oh) * p.out_w + ow) * p.out_c + t.ns * p.N
                                                    : nullptr;
#pragma unroll
                        for (int cc = 0; cc < XC; ++cc) {
                            if (cc < n32) {
                                float v[32];
                                tmem_ld32(t_acc + cc * 32, v);  // GDN: the norm; otherwise the accumulator
                                tmem_ld_wait();
                                if (kGdn) {
                                    if (pre_px && cc * 32 < c_left) {
#pragma unroll
                                        for (int q = 0; q < 4; ++q)
                                            reinterpret_cast<uint4*>(pre_px + cc * 32)[q] =
                                                make_uint4(xs[cc * 16 + 4 * q], xs[cc * 16 + 4 * q + 1], xs[cc * 16 + 4 * q + 2], xs[cc * 16 + 4 * q + 3]);
                                    }
                                    uint32_t out[16];
                                    gdn_stage2_32<EPI == LICOS_EPI_IGDN>(v, beta_s + cc * 32, xs + cc * 16, out);
                                    if (OUT_NHWC) {
                                        store_row32(staging, et, cc, out);
                                    } else if (in_range) {
#pragma unroll
                                        for (int j = 0; j < 16; ++j) {
                                            if (cc * 32 + 2 * j < c_left) o[(size_t)(cc * 32 + 2 * j) * cs] = __uint_as_float(out[j] << 16);
                                            if (cc * 32 + 2 * j + 1 < c_left) o[(size_t)(cc * 32 + 2 * j + 1) * cs] = __uint_as_float(out[j] & 0xffff0000u);
                                        }
                                    }
                                } else {
                                    const float4* b4 = reinterpret_cast<const float4*>(bias_t + cc * 32);
#pragma unroll
                                    for (int q = 0; q < 8; ++q) {
                                        const float4 b = b4[q];
                                        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
                                        if (EPI == LICOS_EPI_RELU) {
#pragma unroll
                                            for (int i = 0; i < 4; ++i) v[4 * q + i] = fmaxf(v[4 * q + i], 0.f);
                                        }
                                    }
                                    if (OUT_NHWC) {
                                        uint32_t out[16];
#pragma unroll
                                        for (int i = 0; i < 16; ++i) out[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                                        store_row32(staging, et, cc, out);
                                    } else if (in_range) {
#pragma unroll
                                        for (int j = 0; j < 32; ++j)
                                            if (cc * 32 + j < c_left) o[(size_t)(cc * 32 + j) * cs] = v[j];
                                    }
                                }
                            }
                        }
                        if (!OUT_NHWC && !kGdn && (p.N & 16)) {  // trailing 16 columns (N % 32 == 16)
                            float h[16];
                            tmem_ld16(t_acc + n32 * 32, h);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float x = h[j] + bias_t[n32 * 32 + j];
                                if (EPI == LICOS_EPI_RELU) x = fmaxf(x, 0.f);
                                if (in_range && n32 * 32 + j < c_left) o[(size_t)(n32 * 32 + j) * cs] = x;
                            }
                        }
                        // this accumulator has been read: hand it back to the MMA issuer
                        tc_fence_before();
                        mbar_arrive(&acc_empty[buf]);
                        if (probe) e_s2 += clock64() - _s2;
                        if (OUT_NHWC) {
                            const long long _st = probe ? clock64() : 0;
                            fence_proxy_async();
                            named_bar_sync(1 + team, 128);
                            if (leader) {
                                for (int at = 0; at < p.N / kKChunk; ++at) {
                                    tma_store_4d(&p.out_maps[ps.out_map[g]], staging + (size_t)at * (128 * 128),
                                                 at * kKChunk, t.gw0, t.gh0 + a * kAccRows, t.b);
                                }
                                tma_store_commit();
                            }
                            if (probe) e_store += clock64() - _st;
                        }
                    }
                }
            }
        }
        if (leader) tma_store_wait_all();
        if (probe) {
            unsigned long long* d = p.dbg + blockIdx.x * 16;
            d[DBG_EPI_ACC] = e_acc; d[DBG_EPI_S1] = e_s1; d[DBG_EPI_NORM] = e_norm; d[DBG_EPI_S2] = e_s2;
            d[DBG_EPI_STORE] = e_store; d[DBG_EPI_TOTAL] = clock64() - t_begin;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace licos

#include "conv_pair.cuh"

namespace licos {

// ----------------------------------------------------------------------------------------------
// small helper kernels: weight / GDN packing, first-layer im2col
// ----------------------------------------------------------------------------------------------

// packed[t][o][i], t = kh*KW + kw
__global__ void pack_weight_kernel(const float* __restrict__ w, int transposed, int out_c, int in_c, int KH, int KW,
                                   int rows, int cin_pad, __nv_bfloat16* __restrict__ packed) {
    const int64_t total = (int64_t)KH * KW * rows * cin_pad;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int i = (int)(e % cin_pad);
        const int o = (int)((e / cin_pad) % rows);
        const int t = (int)(e / ((int64_t)cin_pad * rows));
        float v = 0.f;
        if (o < out_c && i < in_c) {
            const int kh = t / KW, kw = t % KW;
            v = transposed ? w[(((size_t)i * out_c + o) * KH + kh) * KW + kw]
                           : w[(((size_t)o * in_c + i) * KH + kh) * KW + kw];
        }
        packed[e] = __float2bfloat16_rn(v);
    }
}

// first layer: packed[o][k], k = (c*KH + kh)*KW + kw == the flattened torch weight row, zero padded
__global__ void pack_weight_first_kernel(const float* __restrict__ w, int out_c, int K, int rows, int k_pad,
                                         __nv_bfloat16* __restrict__ packed) {
    const int64_t total = (int64_t)rows * k_pad;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int k = (int)(e % k_pad);
        const int o = (int)(e / k_pad);
        packed[e] = __float2bfloat16_rn((o < out_c && k < K) ? w[(size_t)o * K + k] : 0.f);
    }
}

__global__ void gdn_pack_kernel(const float* __restrict__ beta, const float* __restrict__ gamma, int C, float bb,
                                float gb, float ped, float* __restrict__ beta_hat, __nv_bfloat16* __restrict__ gamma_hat) {
    const int64_t total = (int64_t)C * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const float g = fmaxf(gamma[e], gb);
        gamma_hat[e] = __float2bfloat16_rn(g * g - ped);
        if (e < C) {
            const float b = fmaxf(beta[e], bb);
            beta_hat[e] = b * b - ped;
        }
    }
}

// rows[(b, oh, ow)][k] = x[b][c][2*oh + kh - 2][2*ow + kw - 2], k = (c*5 + kh)*5 + kw; one 16-byte
// group of 8 k's per thread.
__global__ void im2col_first_kernel(const float* __restrict__ x, int B, int C, int H, int W, int OH, int OW, int k_pad,
                                    __nv_bfloat16* __restrict__ rows) {
    const int groups = k_pad / 8;
    const int64_t total = (int64_t)B * OH * OW * groups;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int K = C * 25;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int gk = (int)(e % groups);
        int64_t pix = e / groups;
        const int ow = (int)(pix % OW);
        pix /= OW;
        const int oh = (int)(pix % OH);
        const int b = (int)(pix / OH);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = gk * 8 + j;
            float val = 0.f;
            if (k < K) {
                const int c = k / 25, r = k % 25;
                const int ih = 2 * oh + r / 5 - 2, iw = 2 * ow + r % 5 - 2;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = __ldg(x + (((size_t)b * C + c) * H + ih) * W + iw);
            }
            v[j] = val;
        }
        *reinterpret_cast<uint4*>(rows + (size_t)e * 8) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

// bf16 tensor, up to 4 dims, innermost dim contiguous; strides in elements for dims 1..rank-1
static bool make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                     const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdims[4], gstrides[3];
    cuuint32_t gbox[4], estr[4];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (dims[i] == 0) return false;
    }
    for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_elems[i] * 2;
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides,
                          gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

struct NPlan {
    int N, n_split, rows;
};
static NPlan plan_n(int out_c) {
    NPlan pl;
    const int pad = (out_c + 15) / 16 * 16;
    pl.n_split = (pad + 255) / 256;
    pl.N = ((pad + pl.n_split - 1) / pl.n_split + 15) / 16 * 16;
    pl.rows = pl.N * pl.n_split;
    return pl;
}
// g_s[6]-style layer: ConvTranspose2d to <= 4 channels -> GEMM + gather kernel
static bool use_narrow(int kind, int out_c, int in_c) { return kind == LICOS_DECONV_5X5_S2 && out_c <= kN2Cpt && in_c % 64 == 0 && in_c <= 256; }
// g_a[0]-style layer: K = 25 * C_in fits two swizzle atoms -> fused in-kernel im2col
static bool use_first_direct(int kind, int in_c, int out_c, int out_layout) {
    return kind == LICOS_CONV_5X5_S2 && in_c * 25 <= 128 && out_layout == LICOS_LAYOUT_NHWC_BF16 && out_c % 64 == 0 &&
           out_c <= 256;
}
static int taps_of(int kind) { return kind == LICOS_CONV_3X3_S1 ? 9 : (kind == LICOS_CONV_1X1 ? 1 : 25); }
static int first_kpad(int in_c) { return (in_c * 25 + 63) / 64 * 64; }

static int ew_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

static unsigned long long* g_conv_probe = nullptr;

// Development knobs, read once (tools/probe_conv.py experiments; none is needed in production):
//   LICOS_FIRST_V1=1       route the first layer to the non-pipelined kernel
//   LICOS_PREFER_NACC2=1   two single-buffered accumulators instead of one double-buffered one when both do not fit
//   LICOS_NO_SMALL_TILES=1 keep 16-row tiles even when there are fewer of them than SMs
//   LICOS_FORCE_NACC1=1    8-row tiles everywhere (tile-quantisation experiments at small batches)
//   LICOS_NO_PAIR=1        single-CTA engine (cta_group::1) everywhere: the A/B switch for the CTA-pair kernel
//   LICOS_NO_WIDE=1        CTA pairs on the per-column-tap slabs (no wide slabs): the A/B switch for the wide-slab geometry
//   LICOS_SA / LICOS_SB    force the slab / weight ring depths
//   LICOS_DBG_FLAGS        bit 0 / 1: load every slab / weight ring slot only once (isolates the mainloop from data movement)
struct DevKnobs {
    bool first_v1, prefer_nacc2, no_small_tiles, force_nacc1, no_pair, no_wide;
    int sa, sb, dbg_flags;
};
static const DevKnobs& knobs() {
    static const DevKnobs k = [] {
        DevKnobs v{};
        v.first_v1 = getenv("LICOS_FIRST_V1") != nullptr;
        v.prefer_nacc2 = getenv("LICOS_PREFER_NACC2") != nullptr;
        v.no_small_tiles = getenv("LICOS_NO_SMALL_TILES") != nullptr;
        v.force_nacc1 = getenv("LICOS_FORCE_NACC1") != nullptr;
        v.no_pair = getenv("LICOS_NO_PAIR") != nullptr;
        v.no_wide = getenv("LICOS_NO_WIDE") != nullptr;
        if (const char* e = getenv("LICOS_SA")) v.sa = atoi(e);
        if (const char* e = getenv("LICOS_SB")) v.sb = atoi(e);
        if (const char* e = getenv("LICOS_DBG_FLAGS")) v.dbg_flags = atoi(e);
        return v;
    }();
    return k;
}

static void set_tap(Slab& s, int i, int row_off, int group, int w_tap) {
    s.taps[i].row_off = (int8_t)row_off;
    s.taps[i].group = (int8_t)group;
    s.taps[i].w_tap = (int16_t)w_tap;
}

static int sm_count_of(const licos_conv_args* a, int* sms) {
    *sms = a->sm_count;
    if (*sms <= 0) {
        int dev = 0;
        LICOS_CUDA_OK(cudaGetDevice(&dev));
        LICOS_CUDA_OK(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    }
    return LICOS_OK;
}

static int launch_first(const licos_conv_args* a, cudaStream_t s) {
    const bool gdn = (a->epilogue == LICOS_EPI_GDN || a->epilogue == LICOS_EPI_IGDN);
    FirstParams p;
    memset(&p, 0, sizeof(p));
    p.x = (const float*)a->in;
    p.bias = a->bias;
    p.beta = a->beta;
    p.B = a->batch; p.C = a->in_c; p.H = a->in_h; p.W = a->in_w;
    p.OH = (a->in_h + 1) / 2; p.OW = (a->in_w + 1) / 2;
    p.N = a->out_c; p.K = a->in_c * 25; p.k_pad = first_kpad(a->in_c);
    p.tiles_h = (p.OH + 7) / 8; p.tiles_w = (p.OW + 15) / 16;
    const int64_t tiles = (int64_t)a->batch * p.tiles_h * p.tiles_w;
    if (tiles > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    p.total_tiles = (int)tiles;
    const int need_cols = gdn ? 2 * p.N : p.N;
    p.tmem_cols = need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u);
    {
        const uint64_t dims[2] = {(uint64_t)p.k_pad, (uint64_t)p.N};
        const uint64_t strides[1] = {(uint64_t)p.k_pad};
        const uint32_t box[2] = {64, (uint32_t)p.N};
        if (!make_map(&p.w_map, a->weight, 2, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    if (gdn) {
        const uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.N};
        const uint64_t strides[1] = {(uint64_t)p.N};
        const uint32_t box[2] = {64, (uint32_t)p.N};
        if (!make_map(&p.g_map, a->gamma, 2, dims, strides, box)) return LICOS_ERR_CUDA;
    } else {
        p.g_map = p.w_map;
    }
    {
        const uint64_t OC = (uint64_t)p.N, OH = (uint64_t)p.OH, OW = (uint64_t)p.OW;
        const uint64_t dims[4] = {OC, OW, OH, (uint64_t)a->batch};
        const uint64_t strides[3] = {OC, OW * OC, OH * OW * OC};
        const uint32_t box[4] = {64, 16, 8, 1};
        if (!make_map(&p.out_map, a->out, 4, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    const int k_atoms = p.k_pad / 64, n_atoms = p.N / 64;
    const int as_atoms = k_atoms > n_atoms ? k_atoms : n_atoms;
    const size_t smem = 1024 + (size_t)k_atoms * p.N * 128 + (gdn ? (size_t)n_atoms * p.N * 128 : 0) +
                        (size_t)as_atoms * 128 * 128 + (size_t)p.C * kPatchRows * kPatchPitch * 4;
    if (smem > (size_t)kMaxDynSmem) return LICOS_ERR_UNSUPPORTED;
    int sms = 0;
    const int rc = sm_count_of(a, &sms);
    if (rc != LICOS_OK) return rc;
    const int per_sm = (smem <= 113000 && p.tmem_cols <= 256) ? 2 : 1;
    const int grid = (int)(tiles < (int64_t)sms * per_sm ? tiles : (int64_t)sms * per_sm);
    cudaError_t err = cudaErrorInvalidValue;
#define LICOS_LAUNCH_FIRST(E)                                                                                       \
    do {                                                                                                            \
        const cudaError_t attr = ensure_max_dynamic_smem((const void*)conv_first_kernel<E>, kMaxDynSmem);          \
        if (attr != cudaSuccess) { err = attr; break; }                                                             \
        conv_first_kernel<E><<<grid, kEdgeThreads, smem, s>>>(p);                                                   \
        err = cudaGetLastError();                                                                                   \
    } while (0)
    switch (a->epilogue) {
        case LICOS_EPI_NONE: LICOS_LAUNCH_FIRST(LICOS_EPI_NONE); break;
        case LICOS_EPI_GDN: LICOS_LAUNCH_FIRST(LICOS_EPI_GDN); break;
        case LICOS_EPI_IGDN: LICOS_LAUNCH_FIRST(LICOS_EPI_IGDN); break;
        case LICOS_EPI_RELU: LICOS_LAUNCH_FIRST(LICOS_EPI_RELU); break;
        default: return LICOS_ERR_INVALID;
    }
#undef LICOS_LAUNCH_FIRST
    LICOS_CUDA_OK(err);
    return LICOS_OK;
}

// fp32 / u8 / u16 tensor, 4 dims, innermost contiguous, no swizzle (input patches of the fused first layer)
static bool make_map_plain(CUtensorMap* m, const void* base, int elem_bytes, const uint64_t* dims,
                           const uint64_t* strides_elems, const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdims[4], gstrides[3];
    cuuint32_t gbox[4], estr[4];
    for (int i = 0; i < 4; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i < 3; ++i) gstrides[i] = strides_elems[i] * (uint64_t)elem_bytes;
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                   : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    const CUresult r = fn(m, dt, 4, const_cast<void*>(base), gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static bool is_pixel_layout(int layout) {
    return layout == LICOS_LAYOUT_NCHW_F32 || layout == LICOS_LAYOUT_NCHW_U8 || layout == LICOS_LAYOUT_NCHW_U16 ||
           layout == LICOS_LAYOUT_NCHW_U16_Q8;
}
static int first2_mode_of(int layout) {
    return layout == LICOS_LAYOUT_NCHW_U8 ? kF2InU8
                                          : (layout == LICOS_LAYOUT_NCHW_U16 ? kF2InU16 : (layout == LICOS_LAYOUT_NCHW_U16_Q8 ? kF2InU16Q8 : kF2InF32));
}
static int default_int_max(int layout) { return layout == LICOS_LAYOUT_NCHW_U8 ? 255 : 4095; }

// bf16 bits of an fp32 value, round to nearest even (finite inputs)
static uint16_t host_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
// The builders turn an integer pixel v into bf16(float(v) * fl(1 / int_max)) (one FMUL).  The fp32 path feeds
// bf16(float32(float64(v) / int_max)) (raw_image_folder.py:192 followed by the tensor conversion at :196).  The two agree
// for every v of every int_max tried (255, 1023, 4095, 16383, 65535), but that is a property of the constant, so it is
// CHECKED here, exhaustively, the first time an int_max is used; the same for the exact integer form of the 8-bit
// re-quantisation.  A failing int_max is refused (LICOS_ERR_UNSUPPORTED), never approximated.
static bool int_input_exact(int int_max, bool q8, uint32_t* magic_out) {
    static std::mutex mu;
    static int ok_plain[8], ok_q8[8], n_plain = 0, n_q8 = 0;
    if (int_max < 1 || int_max > 65535) return false;
    const uint32_t magic = (uint32_t)((((uint64_t)1 << 43) + 2ull * int_max - 1) / (2ull * int_max));
    if (magic_out) *magic_out = magic;
    std::lock_guard<std::mutex> lock(mu);
    int* list = q8 ? ok_q8 : ok_plain;
    int& n = q8 ? n_q8 : n_plain;
    for (int i = 0; i < n; ++i)
        if (list[i] == int_max) return true;
    if (q8) {
        if (int_max % 2 == 0) return false;  // an even int_max has exact ties in v / int_max * 255
        const float s255 = (float)(1.0 / 255.0);
        for (int v = 0; v <= int_max; ++v) {
            const double t = (double)v / int_max * 255.0;
            double r = nearbyint(t);
            if (r < 0) r = 0;
            if (r > 255) r = 255;
            const uint32_t q = (uint32_t)((((uint64_t)((uint32_t)v * 510u + (uint32_t)int_max)) * magic) >> 43);
            if (q != (uint32_t)r) return false;
            if (host_bf16((float)q * s255) != host_bf16((float)((double)q / 255.0))) return false;
        }
    } else {
        const float sc = (float)(1.0 / (double)int_max);
        for (int v = 0; v <= int_max; ++v)
            if (host_bf16((float)v * sc) != host_bf16((float)((double)v / (double)int_max))) return false;
    }
    if (n < 8) list[n++] = int_max;
    return true;
}

// pipelined first layer (conv_first2.cuh): 1 or 3 bands, N in {64, 128, 192}, rows of x 16-byte aligned (TMA)
static bool use_first2(const licos_conv_args* a) {
    const int eb = first2_elem_bytes(first2_mode_of(a->in_layout));
    return (a->in_c == 1 || a->in_c == 3) && (a->out_c == 64 || a->out_c == 128 || a->out_c == 192) &&
           ((int64_t)a->in_w * eb) % 16 == 0 && ((uintptr_t)a->in & 15) == 0 &&
           (!knobs().first_v1 || a->in_layout != LICOS_LAYOUT_NCHW_F32);
}

static int launch_first2(const licos_conv_args* a, cudaStream_t s) {
    const bool gdn = (a->epilogue == LICOS_EPI_GDN || a->epilogue == LICOS_EPI_IGDN);
    First2Params p;
    memset(&p, 0, sizeof(p));
    p.w = (const __nv_bfloat16*)a->weight;
    p.bias = a->bias;
    p.beta = a->beta;
    p.N = a->out_c;
    p.k_pad = first_kpad(a->in_c);
    const int OH = (a->in_h + 1) / 2, OW = (a->in_w + 1) / 2;
    p.pre_out = (__nv_bfloat16*)a->pre_act;
    p.out_h = OH; p.out_w = OW;
    p.tiles_h = (OH + 7) / 8; p.tiles_w = (OW + 15) / 16;
    const int64_t tiles = (int64_t)a->batch * p.tiles_h * p.tiles_w;
    if (tiles > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    p.total_tiles = (int)tiles;
    const int teams = p.N <= 128 ? 2 : 1;
    p.tmem_cols = 512u;
    if (p.N == 64) p.tmem_cols = 256u;
    p.in_mode = first2_mode_of(a->in_layout);
    if (p.in_mode != kF2InF32) {
        const int int_max = a->int_max > 0 ? a->int_max : default_int_max(a->in_layout);
        if (p.in_mode == kF2InU8 && int_max > 255) return LICOS_ERR_INVALID;
        if (!int_input_exact(int_max, false, &p.q8_magic)) return LICOS_ERR_UNSUPPORTED;
        p.in_scale = (float)(1.0 / (double)int_max);
        p.q8_add = (uint32_t)int_max;
        if (p.in_mode == kF2InU16Q8) {
            if (!int_input_exact(int_max, true, &p.q8_magic)) return LICOS_ERR_UNSUPPORTED;
            p.in_scale = (float)(1.0 / 255.0);
        }
    }
    {
        const uint64_t W = (uint64_t)a->in_w, H = (uint64_t)a->in_h, C = (uint64_t)a->in_c;
        const uint64_t dims[4] = {W, H, C, (uint64_t)a->batch};
        const uint64_t strides[3] = {W, H * W, C * H * W};
        const uint32_t box[4] = {(uint32_t)first2_pitch(p.in_mode), (uint32_t)kF2PatchRows, (uint32_t)a->in_c, 1};
        if (!make_map_plain(&p.x_map, a->in, first2_elem_bytes(p.in_mode), dims, strides, box)) return LICOS_ERR_CUDA;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p.k_pad, (uint64_t)p.N};
        const uint64_t strides[1] = {(uint64_t)p.k_pad};
        const uint32_t box[2] = {64, (uint32_t)p.N};
        if (!make_map(&p.w_map, a->weight, 2, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    if (gdn) {
        const uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.N};
        const uint64_t strides[1] = {(uint64_t)p.N};
        const uint32_t box[2] = {64, (uint32_t)p.N};
        if (!make_map(&p.g_map, a->gamma, 2, dims, strides, box)) return LICOS_ERR_CUDA;
    } else {
        p.g_map = p.w_map;
    }
    {
        const uint64_t OC = (uint64_t)p.N;
        const uint64_t dims[4] = {OC, (uint64_t)OW, (uint64_t)OH, (uint64_t)a->batch};
        const uint64_t strides[3] = {OC, OW * OC, (uint64_t)OH * OW * OC};
        const uint32_t box[4] = {64, 16, 8, 1};
        if (!make_map(&p.out_map, a->out, 4, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    size_t smem = first2_smem_bytes(a->in_c, p.N, gdn, teams);
    if (smem > (size_t)kMaxDynSmem) return LICOS_ERR_UNSUPPORTED;
    if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM (TMEM is sized for that)
    int sms = 0;
    const int rc = sm_count_of(a, &sms);
    if (rc != LICOS_OK) return rc;
    const int grid = (int)(tiles < sms ? tiles : sms);
    cudaError_t err = cudaErrorInvalidValue;
#define LICOS_LAUNCH_F2(E, C, T)                                                                                        \
    do {                                                                                                                \
        const cudaError_t attr = ensure_max_dynamic_smem((const void*)conv_first2_kernel<E, C, T>, kMaxDynSmem);       \
        if (attr != cudaSuccess) { err = attr; break; }                                                                 \
        conv_first2_kernel<E, C, T><<<grid, first2_threads(T), smem, s>>>(p);                                           \
        err = cudaGetLastError();                                                                                       \
    } while (0)
#define LICOS_LAUNCH_F2E(E)                                                          \
    do {                                                                             \
        if (teams == 2) {                                                            \
            if (a->in_c == 1) LICOS_LAUNCH_F2(E, 1, 2); else LICOS_LAUNCH_F2(E, 3, 2); \
        } else {                                                                     \
            if (a->in_c == 1) LICOS_LAUNCH_F2(E, 1, 1); else LICOS_LAUNCH_F2(E, 3, 1); \
        }                                                                            \
    } while (0)
    switch (a->epilogue) {
        case LICOS_EPI_NONE: LICOS_LAUNCH_F2E(LICOS_EPI_NONE); break;
        case LICOS_EPI_GDN: LICOS_LAUNCH_F2E(LICOS_EPI_GDN); break;
        case LICOS_EPI_IGDN: LICOS_LAUNCH_F2E(LICOS_EPI_IGDN); break;
        case LICOS_EPI_RELU: LICOS_LAUNCH_F2E(LICOS_EPI_RELU); break;
        default: return LICOS_ERR_INVALID;
    }
#undef LICOS_LAUNCH_F2E
#undef LICOS_LAUNCH_F2
    LICOS_CUDA_OK(err);
    return LICOS_OK;
}

static int launch_narrow(const licos_conv_args* a, cudaStream_t s) {
    Narrow2Params p;
    memset(&p, 0, sizeof(p));
    p.out = (float*)a->out;
    p.bias = a->bias;
    p.B = a->batch; p.H = a->in_h; p.W = a->in_w; p.C = a->in_c; p.out_c = a->out_c;
    p.OH = 2 * a->in_h; p.OW = 2 * a->in_w;
    p.chunks = a->in_c / 64;
    p.relu = a->epilogue == LICOS_EPI_RELU;
    p.out_mode = a->out_layout == LICOS_LAYOUT_NCHW_U8 ? 1 : (a->out_layout == LICOS_LAYOUT_NCHW_U16 ? 2 : 0);
    if (p.out_mode) {
        const int int_max = a->int_max > 0 ? a->int_max : default_int_max(a->out_layout);
        if (int_max > (p.out_mode == 1 ? 255 : 65535)) return LICOS_ERR_INVALID;
        p.out_scale = (float)int_max;
    }
    p.strip_rows = a->in_h < 32 ? a->in_h : 32;
    p.strips = (a->in_h + p.strip_rows - 1) / p.strip_rows;
    p.segs = (a->in_w + kN2SegPx - 1) / kN2SegPx;
    const int64_t units = (int64_t)a->batch * p.strips * p.segs;
    if (units > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    p.total_units = (int)units;
    {
        const uint64_t C = (uint64_t)a->in_c, H = (uint64_t)a->in_h, W = (uint64_t)a->in_w;
        const uint64_t dims[4] = {C, W, H, (uint64_t)a->batch};
        const uint64_t strides[3] = {C, W * C, H * W * C};
        const uint32_t box[4] = {64, (uint32_t)kN2BoxPx, 1, 1};
        if (!make_map(&p.in_map, a->in, 4, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    {
        const uint64_t dims[2] = {(uint64_t)a->in_c, (uint64_t)(3 * kN2N)};
        const uint64_t strides[1] = {(uint64_t)a->in_c};
        const uint32_t box[2] = {64, (uint32_t)kN2N};
        if (!make_map(&p.w_map, a->weight, 2, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    const size_t w_bytes = (size_t)3 * p.chunks * kN2WTile, slot_bytes = (size_t)p.chunks * kN2ChunkStride;
    int slots = (int)(((size_t)kMaxDynSmem - 1024 - w_bytes) / slot_bytes);
    if (slots > kN2MaxSlots) slots = kN2MaxSlots;
    if (slots < 2) return LICOS_ERR_UNSUPPORTED;
    p.slots = slots;
    size_t smem = 1024 + w_bytes + (size_t)slots * slot_bytes;
    if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: the accumulator ring owns all of TMEM
    int sms = 0;
    const int rc = sm_count_of(a, &sms);
    if (rc != LICOS_OK) return rc;
    const int grid = (int)(units < sms ? units : sms);
    LICOS_CUDA_OK(ensure_max_dynamic_smem((const void*)deconv_narrow2_kernel, kMaxDynSmem));
    deconv_narrow2_kernel<<<grid, kN2Threads, smem, s>>>(p);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

}  // namespace licos

using namespace licos;

extern "C" {

int licos_device_ok(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LICOS_ERR_NO_DEVICE;
    return prop.major == 10 ? LICOS_OK : LICOS_ERR_NO_DEVICE;
}

int licos_pixel_scale_exact(int int_max, int requant8) {
    return int_input_exact(int_max, requant8 != 0, nullptr) ? 1 : 0;
}

int64_t licos_packed_weight_bytes(int kind, int out_c, int in_c, int in_layout) {
    if (out_c < 1 || in_c < 1) return LICOS_ERR_INVALID;
    const NPlan pl = plan_n(out_c);
    if (is_pixel_layout(in_layout)) {
        if (kind != LICOS_CONV_5X5_S2 || in_c > 16) return LICOS_ERR_UNSUPPORTED;
        return (int64_t)pl.rows * first_kpad(in_c) * 2;
    }
    const int cin_pad = (in_c + 63) / 64 * 64;
    if (use_narrow(kind, out_c, in_c)) return (int64_t)3 * kN2N * cin_pad * 2;
    return (int64_t)taps_of(kind) * pl.rows * cin_pad * 2;
}

int licos_pack_conv_weight(const float* w, int kind, int out_c, int in_c, int in_layout, void* packed, void* stream) {
    if (!w || !packed || out_c < 1 || in_c < 1) return LICOS_ERR_INVALID;
    if (kind != LICOS_CONV_5X5_S2 && kind != LICOS_DECONV_5X5_S2 && kind != LICOS_CONV_3X3_S1 && kind != LICOS_CONV_1X1)
        return LICOS_ERR_INVALID;
    const NPlan pl = plan_n(out_c);
    cudaStream_t s = (cudaStream_t)stream;
    if (is_pixel_layout(in_layout)) {
        if (kind != LICOS_CONV_5X5_S2 || in_c > 16) return LICOS_ERR_UNSUPPORTED;
        const int kp = first_kpad(in_c);
        pack_weight_first_kernel<<<ew_grid((int64_t)pl.rows * kp), 256, 0, s>>>(w, out_c, in_c * 25, pl.rows, kp,
                                                                               (__nv_bfloat16*)packed);
    } else if (use_narrow(kind, out_c, in_c)) {
        const int cin_pad = (in_c + 63) / 64 * 64;
        pack_weight_narrow2_kernel<<<ew_grid((int64_t)3 * kN2N * cin_pad), 256, 0, s>>>(w, out_c, in_c, cin_pad,
                                                                                      (__nv_bfloat16*)packed);
    } else {
        const int cin_pad = (in_c + 63) / 64 * 64;
        const int K = kind == LICOS_CONV_3X3_S1 ? 3 : (kind == LICOS_CONV_1X1 ? 1 : 5);
        pack_weight_kernel<<<ew_grid((int64_t)K * K * pl.rows * cin_pad), 256, 0, s>>>(
            w, kind == LICOS_DECONV_5X5_S2, out_c, in_c, K, K, pl.rows, cin_pad, (__nv_bfloat16*)packed);
    }
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_gdn_pack(const float* beta, const float* gamma, int channels, float beta_bound, float gamma_bound,
                   float pedestal, float* beta_hat, void* gamma_hat_bf16, void* stream) {
    if (!beta || !gamma || !beta_hat || !gamma_hat_bf16 || channels < 1) return LICOS_ERR_INVALID;
    gdn_pack_kernel<<<ew_grid((int64_t)channels * channels), 256, 0, (cudaStream_t)stream>>>(
        beta, gamma, channels, beta_bound, gamma_bound, pedestal, beta_hat, (__nv_bfloat16*)gamma_hat_bf16);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

// Development aid, deliberately NOT part of include/licos_b200.h: when non-NULL, every conv launch writes 16 uint64 cycle
// counters per CTA (waits of each warp role, epilogue stages) to this device buffer (>= 16 * SM count entries).
// tools/probe_conv.py binds it ad hoc.
void licos_internal_set_conv_probe(unsigned long long* device_buf) { g_conv_probe = device_buf; }

int64_t licos_conv_workspace_bytes(const licos_conv_args* a) {
    if (!a) return LICOS_ERR_INVALID;
    if (a->in_layout != LICOS_LAYOUT_NCHW_F32) return 0;
    if (use_first_direct(a->kind, a->in_c, a->out_c, a->out_layout)) return 0;
    const int oh = (a->in_h + 1) / 2, ow = (a->in_w + 1) / 2;
    return (int64_t)a->batch * oh * ow * first_kpad(a->in_c) * 2;
}

// `setmaxnreg.inc` draws on the registers the CTA's own warps released and was launched with: a kernel built with fewer
// registers per thread than the budget assumes would wait for ever, so refuse to launch it instead.
static bool pair_regs_ok(const void* kernel, bool pipelined, int threads) {
#ifdef LICOS_NO_EPI_PIPE
    (void)kernel; (void)pipelined; (void)threads;
    return true;
#else
    if (!pipelined) return true;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return false;
    return (int64_t)fa.numRegs * threads >= 128 * kPairCtrlRegs + (int64_t)(threads - 128) * kPairTeamRegs;
#endif
}

int licos_conv_forward(const licos_conv_args* a, void* stream) {
    if (!a || !a->in || !a->out || !a->weight) return LICOS_ERR_INVALID;
    if (a->batch < 0 || a->in_h < 1 || a->in_w < 1 || a->in_c < 1 || a->out_c < 1) return LICOS_ERR_INVALID;
    if (a->batch == 0) return LICOS_OK;
    const bool gdn = (a->epilogue == LICOS_EPI_GDN || a->epilogue == LICOS_EPI_IGDN);
    if (a->epilogue < LICOS_EPI_NONE || a->epilogue > LICOS_EPI_RELU) return LICOS_ERR_INVALID;
    if (gdn && (!a->beta || !a->gamma || !a->bias)) return LICOS_ERR_INVALID;
    if (a->pre_act && (!gdn || a->out_c % 32 != 0 || ((uintptr_t)a->pre_act & 15) != 0)) return LICOS_ERR_INVALID;
    const bool int_out = a->out_layout == LICOS_LAYOUT_NCHW_U8 || a->out_layout == LICOS_LAYOUT_NCHW_U16;
    if (a->out_layout != LICOS_LAYOUT_NCHW_F32 && a->out_layout != LICOS_LAYOUT_NHWC_BF16 && !int_out) return LICOS_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;

    if (is_pixel_layout(a->in_layout) && a->in_layout != LICOS_LAYOUT_NCHW_F32) {
        // integer tiles: only the pipelined first-layer kernel scales them on the fly
        if (!use_first_direct(a->kind, a->in_c, a->out_c, a->out_layout) || !use_first2(a)) return LICOS_ERR_UNSUPPORTED;
        return launch_first2(a, s);
    }
    if (a->in_layout == LICOS_LAYOUT_NCHW_F32 && use_first_direct(a->kind, a->in_c, a->out_c, a->out_layout))
    {
        if (!use_first2(a) && a->pre_act) return LICOS_ERR_UNSUPPORTED;  // only the pipelined kernel writes the pre-activation
        return use_first2(a) ? launch_first2(a, s) : launch_first(a, s);
    }
    if (a->in_layout == LICOS_LAYOUT_NHWC_BF16 && use_narrow(a->kind, a->out_c, a->in_c)) {
        if ((a->out_layout != LICOS_LAYOUT_NCHW_F32 && !int_out) || gdn) return LICOS_ERR_UNSUPPORTED;
        return launch_narrow(a, s);
    }
    if (int_out) return LICOS_ERR_UNSUPPORTED;  // integer pixels come out of the model's last layer only

    ConvParams p;  // ~3.3 KB of plain data, passed by value as a __grid_constant__ kernel parameter
    memset(&p, 0, sizeof(p));

    const NPlan pl = plan_n(a->out_c);
    if (pl.rows > 512) return LICOS_ERR_UNSUPPORTED;  // bias staging holds 512 channels
    p.N = pl.N;
    p.n_split = pl.n_split;
    p.w_rows_per_tap = pl.rows;
    p.epilogue = a->epilogue;
    p.out_layout = a->out_layout;
    p.out_c = a->out_c;
    p.batch = a->batch;
    p.bias = a->bias;
    p.beta = a->beta;
    p.out_f32 = (a->out_layout == LICOS_LAYOUT_NCHW_F32) ? (float*)a->out : nullptr;
    p.pre_out = (__nv_bfloat16*)a->pre_act;

    if (a->out_layout == LICOS_LAYOUT_NHWC_BF16 && (a->out_c % 64 != 0 || a->out_c > 256)) return LICOS_ERR_UNSUPPORTED;
    if (gdn && (a->out_c % 64 != 0 || a->out_c > 256)) return LICOS_ERR_UNSUPPORTED;

    // ---- geometry of the position grid, the input view(s) and the pass table -------------------
    const void* in_base = a->in;
    int in_h = a->in_h, in_w = a->in_w, cin_pad;
    int kind = a->kind;
    bool pointwise = false;
    if (a->in_layout == LICOS_LAYOUT_NCHW_F32) {
        // first layer: explicit im2col of the tiny-Cin input, then a 1x1 "conv" over K_pad channels
        if (kind != LICOS_CONV_5X5_S2 || a->in_c > 16) return LICOS_ERR_UNSUPPORTED;
        if (!a->workspace || a->workspace_bytes < licos_conv_workspace_bytes(a)) return LICOS_ERR_BUFFER;
        const int oh = (a->in_h + 1) / 2, ow = (a->in_w + 1) / 2;
        const int kp = first_kpad(a->in_c);
        const int64_t groups = (int64_t)a->batch * oh * ow * (kp / 8);
        im2col_first_kernel<<<ew_grid(groups), 256, 0, s>>>((const float*)a->in, a->batch, a->in_c, a->in_h, a->in_w, oh,
                                                            ow, kp, (__nv_bfloat16*)a->workspace);
        LICOS_CUDA_OK(cudaGetLastError());
        in_base = a->workspace;
        in_h = oh;
        in_w = ow;
        cin_pad = kp;
        pointwise = true;
    } else if (a->in_layout == LICOS_LAYOUT_NHWC_BF16) {
        if (a->in_c % 64 != 0) return LICOS_ERR_UNSUPPORTED;
        cin_pad = a->in_c;
    } else {
        return LICOS_ERR_INVALID;
    }
    p.cin_chunks = cin_pad / kKChunk;

    int out_h, out_w;
    if (pointwise) { out_h = in_h; out_w = in_w; p.grid_h = in_h; p.grid_w = in_w; p.out_s = 1; }
    else if (kind == LICOS_CONV_5X5_S2) {
        if (in_h < 2 || in_w < 2) return LICOS_ERR_UNSUPPORTED;
        out_h = (in_h + 1) / 2; out_w = (in_w + 1) / 2; p.grid_h = out_h; p.grid_w = out_w; p.out_s = 1;
    } else if (kind == LICOS_DECONV_5X5_S2) {
        out_h = 2 * in_h; out_w = 2 * in_w; p.grid_h = in_h; p.grid_w = in_w; p.out_s = 2;
    } else if (kind == LICOS_CONV_3X3_S1 || kind == LICOS_CONV_1X1) {
        out_h = in_h; out_w = in_w; p.grid_h = in_h; p.grid_w = in_w; p.out_s = 1;
    } else {
        return LICOS_ERR_INVALID;
    }
    p.out_h = out_h;
    p.out_w = out_w;

    const bool merged = (kind == LICOS_DECONV_5X5_S2 && a->out_layout == LICOS_LAYOUT_NCHW_F32 && pl.N <= 32 &&
                         pl.n_split == 1 && !gdn);
    const int groups = merged ? 4 : 1;

    // CTA pairs (conv_pair.cuh) take every single-group shape; all but the 1x1 layers on wide slabs
    int sms = a->sm_count;
    if (sms <= 0) {
        int dev = 0;
        LICOS_CUDA_OK(cudaGetDevice(&dev));
        LICOS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const bool use_pair = groups == 1 && !knobs().no_pair && sms >= 2 && pl.N % 16 == 0;
    const bool wide = use_pair && !pointwise && kind != LICOS_CONV_1X1 && !knobs().no_wide;

    // accumulators per tile
    // GDN accumulates its norm in place, so TMEM holds only accumulators; two sets when they fit so that the
    // epilogue of one pass overlaps the mainloop of the next
    int n_acc = 2;
    if (groups * 2 * pl.N > (int)kTmemCols) n_acc = 1;
    if (wide ? p.grid_w <= 8 : p.grid_h <= kAccRows) n_acc = 1;
    if (n_acc == 2 && 2 * groups * 2 * pl.N > (int)kTmemCols && 2 * groups * pl.N <= (int)kTmemCols &&
        !knobs().prefer_nacc2)
        n_acc = 1;  // one double-buffered accumulator beats two single-buffered ones that share weight tiles
    if (n_acc == 2 && !knobs().no_small_tiles) {
        // Small batches (a 32-tile training step, the 16 x 16 latent end of the transforms): 16-row tiles would leave SMs
        // without a tile, so take 8-row tiles (twice as many) while those still fit one wave.  Never true at the inference batch.
        int sms_q = 0;
        if (sm_count_of(a, &sms_q) == LICOS_OK) {
            const int64_t tiles16 = (int64_t)a->batch * ((p.grid_h + 2 * kAccRows - 1) / (2 * kAccRows)) *
                                    ((p.grid_w + kTileW - 1) / kTileW) * pl.n_split;
            if (2 * tiles16 <= sms_q) n_acc = 1;  // (measured: with 128 tiles of 148, halving them costs more than it fills)
        }
    }
    if (knobs().force_nacc1) n_acc = 1;  // experiment: 8-row tiles everywhere
    if (groups * n_acc * pl.N > (int)kTmemCols) return LICOS_ERR_UNSUPPORTED;
    p.n_acc = n_acc;
    p.n_buf = (2 * groups * n_acc * pl.N <= (int)kTmemCols) ? 2 : 1;
    p.jobs_per_pass = groups * n_acc;
    const int TH = wide ? 16 : kAccRows * n_acc;       // tile rows
    const int TW = wide ? 8 * n_acc : kTileW;          // tile columns
    const int R = TH + 2;                              // slab rows
    const int SW = wide ? TW + 2 : kTileW;             // slab columns
    p.wide = wide ? 1 : 0;
    p.tile_h = TH;
    p.tile_w = TW;
    p.a_pitch16 = (uint32_t)SW * 8u;
    p.a_sbo16 = wide ? p.a_pitch16 : 64u;
    p.acc_step16 = wide ? 64u : (uint32_t)(kAccRows * kRowBytes) >> 4;

    // ---- passes ------------------------------------------------------------------------------
    int8_t col_offs[kMaxPasses][kMaxSlabs][kMaxTaps];
    memset(col_offs, 0, sizeof(col_offs));
    if (wide) {
        // one slab per parity view / pass, origin one pixel up-left of the tile, every tap a (row, column) offset into it
        auto add_tap = [&](int pi, int si, Slab& sl, int row_off, int col_off, int w_tap) {
            col_offs[pi][si][sl.n_taps] = (int8_t)col_off;
            set_tap(sl, sl.n_taps, row_off, 0, w_tap);
            ++sl.n_taps;
        };
        if (kind == LICOS_CONV_5X5_S2) {
            p.n_passes = 1;
            Pass& ps = p.passes[0];
            ps.n_groups = 1;
            ps.n_slabs = 4;
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw) {
                    const int si = ph * 2 + pw;
                    Slab& sl = ps.slabs[si];
                    sl.in_map = (int8_t)si; sl.dw = -1; sl.n_taps = 0;
                    for (int kh = ph; kh < 5; kh += 2)
                        for (int kw = pw; kw < 5; kw += 2)
                            add_tap(0, si, sl, (kh - 2 - ph) / 2 + 1, (kw - 2 - pw) / 2 + 1, kh * 5 + kw);
                }
        } else if (kind == LICOS_CONV_3X3_S1) {
            p.n_passes = 1;
            Pass& ps = p.passes[0];
            ps.n_groups = 1;
            ps.n_slabs = 1;
            Slab& sl = ps.slabs[0];
            sl.in_map = 0; sl.dw = -1; sl.n_taps = 0;
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw) add_tap(0, 0, sl, kh, kw, kh * 3 + kw);
        } else {  // transposed conv: one pass per output parity, the same input window each time
            p.n_passes = 4;
            for (int pi = 0; pi < 4; ++pi) {
                const int aa = pi >> 1, b = pi & 1;
                Pass& ps = p.passes[pi];
                ps.n_groups = 1;
                ps.dy[0] = (int8_t)aa; ps.dx[0] = (int8_t)b; ps.out_map[0] = (int8_t)pi;
                ps.n_slabs = 1;
                Slab& sl = ps.slabs[0];
                sl.in_map = 0; sl.dw = -1; sl.n_taps = 0;
                for (int kh = aa; kh < 5; kh += 2)
                    for (int kw = b; kw < 5; kw += 2)
                        add_tap(pi, 0, sl, (aa + 2 - kh) / 2 + 1, (b + 2 - kw) / 2 + 1, kh * 5 + kw);
            }
        }
    } else if (pointwise || kind == LICOS_CONV_1X1) {
        p.n_passes = 1;
        Pass& ps = p.passes[0];
        ps.n_slabs = 1; ps.n_groups = 1;
        ps.slabs[0].in_map = 0; ps.slabs[0].dw = 0; ps.slabs[0].n_taps = 1;
        set_tap(ps.slabs[0], 0, 1, 0, 0);
    } else if (kind == LICOS_CONV_5X5_S2) {
        p.n_passes = 1;
        Pass& ps = p.passes[0];
        ps.n_groups = 1;
        int ns = 0;
        for (int kw = 0; kw < 5; ++kw) {
            const int pw = kw & 1, dw = (kw - 2 - pw) / 2;
            for (int ph = 0; ph < 2; ++ph) {
                Slab& sl = ps.slabs[ns++];
                sl.in_map = (int8_t)(ph * 2 + pw);
                sl.dw = (int8_t)dw;
                int nt = 0;
                for (int kh = ph; kh < 5; kh += 2) {
                    const int dh = (kh - 2 - ph) / 2;
                    set_tap(sl, nt++, dh + 1, 0, kh * 5 + kw);
                }
                sl.n_taps = (int8_t)nt;
            }
        }
        ps.n_slabs = (int8_t)ns;
    } else if (kind == LICOS_CONV_3X3_S1) {
        p.n_passes = 1;
        Pass& ps = p.passes[0];
        ps.n_groups = 1;
        ps.n_slabs = 3;
        for (int kw = 0; kw < 3; ++kw) {
            Slab& sl = ps.slabs[kw];
            sl.in_map = 0; sl.dw = (int8_t)(kw - 1); sl.n_taps = 3;
            for (int kh = 0; kh < 3; ++kh) set_tap(sl, kh, kh, 0, kh * 3 + kw);
        }
    } else if (merged) {
        // all four output parities at once: slab per column shift, every tap that uses it
        p.n_passes = 1;
        Pass& ps = p.passes[0];
        ps.n_groups = 4;
        for (int g = 0; g < 4; ++g) { ps.dy[g] = (int8_t)(g >> 1); ps.dx[g] = (int8_t)(g & 1); ps.out_map[g] = (int8_t)g; }
        int ns = 0;
        for (int dw = 1; dw >= -1; --dw) {
            Slab& sl = ps.slabs[ns++];
            sl.in_map = 0; sl.dw = (int8_t)dw;
            int nt = 0;
            for (int kw = 0; kw < 5; ++kw) {
                const int b = kw & 1;
                if ((b + 2 - kw) / 2 != dw) continue;
                for (int kh = 0; kh < 5; ++kh) {
                    const int aa = kh & 1, dh = (aa + 2 - kh) / 2;
                    set_tap(sl, nt++, dh + 1, aa * 2 + b, kh * 5 + kw);
                }
            }
            sl.n_taps = (int8_t)nt;
        }
        ps.n_slabs = (int8_t)ns;
    } else {  // deconv, one pass per output parity
        p.n_passes = 4;
        for (int pi = 0; pi < 4; ++pi) {
            const int aa = pi >> 1, b = pi & 1;
            Pass& ps = p.passes[pi];
            ps.n_groups = 1;
            ps.dy[0] = (int8_t)aa; ps.dx[0] = (int8_t)b; ps.out_map[0] = (int8_t)pi;
            int ns = 0;
            for (int kw = b; kw < 5; kw += 2) {
                Slab& sl = ps.slabs[ns++];
                sl.in_map = 0; sl.dw = (int8_t)((b + 2 - kw) / 2);
                int nt = 0;
                for (int kh = aa; kh < 5; kh += 2) set_tap(sl, nt++, (aa + 2 - kh) / 2 + 1, 0, kh * 5 + kw);
                sl.n_taps = (int8_t)nt;
            }
            ps.n_slabs = (int8_t)ns;
        }
    }

    // ---- lean issue-loop encoding: the taps of every slab must walk consecutive slab rows, one group ----
    p.lean = groups == 1;
    for (int pi = 0; pi < p.n_passes && p.lean; ++pi) {
        const Pass& ps = p.passes[pi];
        unsigned long long info = 0;
        for (int sidx = 0; sidx < ps.n_slabs && p.lean; ++sidx) {
            const Slab& sl = ps.slabs[sidx];
            const int nt = sl.n_taps, r0 = sl.taps[0].row_off;
            const int step = nt > 1 ? sl.taps[1].row_off - r0 : 1;
            if (nt < 1 || nt > 4 || r0 < 0 || r0 > 3 || (step != 1 && step != -1)) p.lean = 0;
            for (int k = 1; k < nt; ++k)
                if (sl.taps[k].row_off != r0 + k * step || sl.taps[k].group != 0) p.lean = 0;
            info |= (unsigned long long)((nt - 1) | (r0 << 2) | ((step < 0 ? 1 : 0) << 4)) << (5 * sidx);
        }
        p.pass_info[pi] = info;
    }

    // ---- CTA pairs (conv_pair.cuh): the taps of every slab; the weight and gamma tiles are split between the two CTAs ----
    if (use_pair) {
        for (int pi = 0; pi < p.n_passes; ++pi)
            for (int si = 0; si < p.passes[pi].n_slabs; ++si) {
                const Slab& sl = p.passes[pi].slabs[si];
                unsigned long long v = (unsigned long long)sl.n_taps;
                for (int k = 0; k < sl.n_taps; ++k)
                    v |= (unsigned long long)((sl.taps[k].row_off & 3) | ((col_offs[pi][si][k] & 3) << 2)) << (4 + 4 * k);
                p.tap_list[pi][si] = v;
            }
    }
    const uint32_t w_box_rows = use_pair ? (uint32_t)pl.N / 2 : (uint32_t)pl.N;

    // ---- tensor maps ---------------------------------------------------------------------------
    const uint64_t C = (uint64_t)cin_pad, H = (uint64_t)in_h, W = (uint64_t)in_w, B = (uint64_t)a->batch;
    const uint32_t in_box[4] = {(uint32_t)kKChunk, (uint32_t)SW, (uint32_t)R, 1};
    if (!pointwise && kind == LICOS_CONV_5X5_S2) {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const uint64_t dims[4] = {C, (W - pw + 1) / 2, (H - ph + 1) / 2, B};
                const uint64_t strides[3] = {2 * C, 2 * W * C, H * W * C};
                const __nv_bfloat16* base = (const __nv_bfloat16*)in_base + ((size_t)ph * W + pw) * C;
                if (!make_map(&p.in_maps[ph * 2 + pw], base, 4, dims, strides, in_box)) return LICOS_ERR_CUDA;
            }
    } else {
        const uint64_t dims[4] = {C, W, H, B};
        const uint64_t strides[3] = {C, W * C, H * W * C};
        if (!make_map(&p.in_maps[0], in_base, 4, dims, strides, in_box)) return LICOS_ERR_CUDA;
        for (int i = 1; i < 4; ++i) p.in_maps[i] = p.in_maps[0];
    }
    {
        const uint64_t w_taps = pointwise ? 1 : (uint64_t)taps_of(kind);
        const uint64_t dims[2] = {C, w_taps * (uint64_t)pl.rows};
        const uint64_t strides[1] = {C};
        const uint32_t box[2] = {(uint32_t)kKChunk, w_box_rows};
        if (!make_map(&p.w_map, a->weight, 2, dims, strides, box)) return LICOS_ERR_CUDA;
    }
    if (gdn) {
        const uint64_t dims[2] = {(uint64_t)a->out_c, (uint64_t)a->out_c};
        const uint64_t strides[1] = {(uint64_t)a->out_c};
        const uint32_t box[2] = {(uint32_t)kKChunk, w_box_rows};
        if (!make_map(&p.g_map, a->gamma, 2, dims, strides, box)) return LICOS_ERR_CUDA;
    } else {
        p.g_map = p.w_map;
    }
    if (a->out_layout == LICOS_LAYOUT_NHWC_BF16) {
        const uint64_t OC = (uint64_t)a->out_c, OH = (uint64_t)out_h, OW = (uint64_t)out_w;
        const uint32_t box[4] = {(uint32_t)kKChunk, wide ? 8u : (uint32_t)kTileW, wide ? 16u : (uint32_t)kAccRows, 1};
        if (kind == LICOS_DECONV_5X5_S2 && !pointwise) {
            for (int pi = 0; pi < 4; ++pi) {
                const int aa = pi >> 1, b = pi & 1;
                const uint64_t dims[4] = {OC, W, H, B};
                const uint64_t strides[3] = {2 * OC, 2 * OW * OC, OH * OW * OC};
                __nv_bfloat16* base = (__nv_bfloat16*)a->out + ((size_t)aa * OW + b) * OC;
                if (!make_map(&p.out_maps[pi], base, 4, dims, strides, box)) return LICOS_ERR_CUDA;
            }
        } else {
            const uint64_t dims[4] = {OC, OW, OH, B};
            const uint64_t strides[3] = {OC, OW * OC, OH * OW * OC};
            if (!make_map(&p.out_maps[0], a->out, 4, dims, strides, box)) return LICOS_ERR_CUDA;
            for (int i = 1; i < 4; ++i) p.out_maps[i] = p.out_maps[0];
        }
    } else {
        for (int i = 0; i < 4; ++i) p.out_maps[i] = p.in_maps[0];
    }

    // ---- shared memory plan --------------------------------------------------------------------
    p.a_tx_bytes = (uint32_t)R * (uint32_t)SW * 128u;
    p.a_slot_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
    p.b_slot_bytes = w_box_rows * 128u;
    p.staging_bytes = (gdn || a->out_layout == LICOS_LAYOUT_NHWC_BF16) ? (uint32_t)(pl.N / kKChunk) * 128u * 128u : 0u;
    p.gamma_bytes = gdn ? (uint32_t)(pl.N / kKChunk) * w_box_rows * 128u : 0u;
    const int64_t b_slot_al = p.b_slot_bytes;  // N is a multiple of 16, so N*128 is a multiple of 2 KB
    int sa = 0, sb = 0;
    for (p.n_teams = 2; p.n_teams >= 1; --p.n_teams) {
        const int64_t budget = kMaxDynSmem - 1024 - (int64_t)p.n_teams * p.staging_bytes - (int64_t)p.gamma_bytes;
        sb = 5;
        sa = 0;
        while (sa < 2 && sb > 2) {  // prefer >= 4 weight slots, give them up when the resident gamma leaves no room
            --sb;
            sa = (int)((budget - (int64_t)sb * b_slot_al) / p.a_slot_bytes);
        }
        if (sa < 2) continue;
        if (sa > kMaxSA) sa = kMaxSA;
        int64_t left = budget - (int64_t)sa * p.a_slot_bytes - (int64_t)sb * b_slot_al;
        while (sb < kMaxSB && left >= b_slot_al) { ++sb; left -= b_slot_al; }
        break;
    }
    if (p.n_teams < 1) return LICOS_ERR_UNSUPPORTED;
    if (knobs().sa >= 2 && knobs().sa <= kMaxSA) sa = knobs().sa;
    if (knobs().sb >= 2 && knobs().sb <= kMaxSB) sb = knobs().sb;
    p.sa = sa;
    p.sb = sb;
    size_t smem_bytes = 1024 + (size_t)sa * p.a_slot_bytes + (size_t)sb * b_slot_al + (size_t)p.n_teams * p.staging_bytes + p.gamma_bytes;
    if (smem_bytes > (size_t)kMaxDynSmem) return LICOS_ERR_UNSUPPORTED;
    if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns

    p.tiles_h = (p.grid_h + TH - 1) / TH;
    p.tiles_w = (p.grid_w + TW - 1) / TW;
    const int64_t tiles = (int64_t)a->batch * p.tiles_h * p.tiles_w * p.n_split;
    if (tiles > 0x7fffffff) return LICOS_ERR_UNSUPPORTED;
    p.total_tiles = (int)tiles;
    p.dbg = g_conv_probe;
    p.dbg_flags = knobs().dbg_flags;

    int grid = (int)(tiles < sms ? tiles : sms);
    if (use_pair) {
        // pair work items: two spatial tiles (one per CTA) that share a channel split
        const int64_t spatial = (int64_t)a->batch * p.tiles_h * p.tiles_w;
        const int64_t items = (spatial + 1) / 2 * p.n_split;
        p.total_tiles = (int)items;
        const int64_t pairs = items < sms / 2 ? items : sms / 2;
        grid = (int)(2 * pairs);
    }
    const bool nhwc = a->out_layout == LICOS_LAYOUT_NHWC_BF16;
    cudaError_t err = cudaErrorInvalidValue;
#define LICOS_LAUNCH_PAIR(E, O, XP, TWP)                                                                \
    do {                                                                                                \
        const cudaError_t attr = ensure_max_dynamic_smem((const void*)conv_igemm_pair_kernel<E, O, XP, TWP>, kMaxDynSmem); \
        if (attr != cudaSuccess) { err = attr; break; }                                                 \
        if (!pair_regs_ok((const void*)conv_igemm_pair_kernel<E, O, XP, TWP>,                           \
                          (E == LICOS_EPI_GDN || E == LICOS_EPI_IGDN) && TWP == 4 && XP <= 4, pair_threads(TWP))) { \
            err = cudaErrorLaunchOutOfResources;                                                        \
            break;                                                                                      \
        }                                                                                               \
        cudaLaunchConfig_t cfg = {};                                                                    \
        cfg.gridDim = dim3((unsigned)grid);                                                             \
        cfg.blockDim = dim3(pair_threads(TWP));                                                         \
        cfg.dynamicSmemBytes = smem_bytes;                                                              \
        cfg.stream = s;                                                                                 \
        cudaLaunchAttribute at[1];                                                                      \
        at[0].id = cudaLaunchAttributeClusterDimension;                                                 \
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;             \
        cfg.attrs = at;                                                                                 \
        cfg.numAttrs = 1;                                                                               \
        err = cudaLaunchKernelEx(&cfg, conv_igemm_pair_kernel<E, O, XP, TWP>, p);                       \
    } while (0)
#define LICOS_LAUNCH_X(E, O, X)                                                                         \
    do {                                                                                                \
        if (use_pair) {                                                                                 \
            LICOS_LAUNCH_PAIR(E, O, X, 4);                                                              \
            break;                                                                                      \
        }                                                                                               \
        const cudaError_t attr = ensure_max_dynamic_smem((const void*)conv_igemm_kernel<E, O, X>,      \
                                                         kMaxDynSmem);                                  \
        if (attr != cudaSuccess) { err = attr; break; }                                                 \
        conv_igemm_kernel<E, O, X><<<grid, kThreads, smem_bytes, s>>>(p);                               \
        err = cudaGetLastError();                                                                       \
    } while (0)
#define LICOS_LAUNCH(E, O)                                                                              \
    do {                                                                                                \
        if (pl.N <= 128) LICOS_LAUNCH_X(E, O, 4);                                                       \
        else if (pl.N <= 192) LICOS_LAUNCH_X(E, O, 6);                                                  \
        else LICOS_LAUNCH_X(E, O, 8);                                                                   \
    } while (0)
    switch (a->epilogue * 2 + (nhwc ? 1 : 0)) {
        case LICOS_EPI_NONE * 2 + 0: LICOS_LAUNCH(LICOS_EPI_NONE, false); break;
        case LICOS_EPI_NONE * 2 + 1: LICOS_LAUNCH(LICOS_EPI_NONE, true); break;
        case LICOS_EPI_GDN * 2 + 0: LICOS_LAUNCH(LICOS_EPI_GDN, false); break;
        case LICOS_EPI_GDN * 2 + 1: LICOS_LAUNCH(LICOS_EPI_GDN, true); break;
        case LICOS_EPI_IGDN * 2 + 0: LICOS_LAUNCH(LICOS_EPI_IGDN, false); break;
        case LICOS_EPI_IGDN * 2 + 1: LICOS_LAUNCH(LICOS_EPI_IGDN, true); break;
        case LICOS_EPI_RELU * 2 + 0: LICOS_LAUNCH(LICOS_EPI_RELU, false); break;
        case LICOS_EPI_RELU * 2 + 1: LICOS_LAUNCH(LICOS_EPI_RELU, true); break;
        default: return LICOS_ERR_INVALID;
    }
#undef LICOS_LAUNCH
#undef LICOS_LAUNCH_X
#undef LICOS_LAUNCH_PAIR
    LICOS_CUDA_OK(err);
    return LICOS_OK;
}

}  // extern "C"
