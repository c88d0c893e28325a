// The conv engine on CTA PAIRS (tcgen05 cta_group::2): the same slab / tap / pass walk, tiles, rings, TMEM double
// buffering and in-place GDN epilogue as conv_igemm_kernel (conv_engine.cu), but two CTAs on the two SMs of a TPC run every
// MMA together with M = 256: each CTA brings its OWN 16 x 16-position tile (its own activation slabs, its own two 128-row
// accumulators in its own TMEM, its own epilogue), the pair shares every weight tile -- each CTA loads and holds only
// HALF of it (N / 2 rows), the tensor cores exchange the halves.
//
// Why (tools/mma_bench7.cu, all 148 SMs busy): a cta_group::1 MMA of M128 N128 K16 reads 4 KB of A + 4 KB of B per 64-cycle
// instruction -- the entire 128 B/clk shared-memory read port -- and runs at 74.5 cycles, 91 with one barrier wait per tap;
// the engine measured 104-110 cycles per MMA with or without its TMA traffic (tools/probe_sweep.sh), i.e. it was never
// bandwidth-bound from L2, it was starved at the operand port.  The pair reads 4 + 2 KB per SM and instruction: 64.0
// cycles per MMA, barrier waits hidden behind the queue.  The fused gamma GEMM runs pair-wide for the same reason (and
// because a kernel must not mix cta_group::1 and ::2 tcgen05 instructions).
//
// Protocol (L = the barrier of the pair's leader, CTA rank 0, is the one that is used; B = each CTA uses its own copy):
//   a_full / b_full [L]   count 1: the leader's producer arrives with expect_tx of BOTH CTAs' bytes; each CTA's TMA load
//                         (cp.async.bulk.tensor ... cta_group::2) lands in its own shared memory and completes its bytes on
//                         the leader's barrier
//   a_empty / b_empty [B] tcgen05.commit ... multicast to both CTAs: each producer refills its own ring
//   acc_full [B]          multicast commit: both epilogues start
//   acc_empty [L]         count 2 x 128 x jobs: every epilogue thread of both CTAs arrives (the peer's remotely)
//   stg_full [L]          count 2 per team: "this CTA's x^2 tile is in shared memory"; then the leader's team warp issues the
//                         pair-wide gamma GEMM and commits (multicast) to norm_full [B]
// Only lean passes (every slab's taps walk consecutive slab rows, one group) take this kernel; the merged narrow deconv
// and other table-walk shapes stay on conv_igemm_kernel.
#pragma once

namespace licos {

struct PairTile {
    int b, gh0, gw0, ns;
    bool live;
};
// work item w of the pair -> this CTA's tile: spatial tile 2 q + rank, channel split ns (the pair shares ns: same weights)
__device__ __forceinline__ PairTile decode_pair_tile(const ConvParams& p, int w, uint32_t rank) {
    PairTile t;
    t.ns = w % p.n_split;
    int sp = (w / p.n_split) * 2 + (int)rank;
    const int spatial = p.batch * p.tiles_h * p.tiles_w;
    t.live = sp < spatial;
    if (!t.live) sp = spatial;  // -> b == batch: every TMA load is zero fill, every TMA store is clipped away
    t.gw0 = (sp % p.tiles_w) * p.tile_w;
    sp /= p.tiles_w;
    t.gh0 = (sp % p.tiles_h) * p.tile_h;
    t.b = sp / p.tiles_h;
    return t;
}

// Epilogue teams of TW warps.  TW = 4 is what is instantiated.  TW = 8 (two warps per TMEM lane quadrant taking alternate
// 32-channel chunks, 256 threads per job) halves the time of the per-row loops but leaves only 96 registers per thread
// (640 threads) and did not shorten a job: measured on the transposed-conv + IGDN layers (8 epilogue jobs per tile), the
// kernel as a whole is bound by the 128 B/clk shared-memory port once the MMAs run on CTA pairs -- operand reads 6 KB per
// MMA, slab / weight TMA writes, the x^2 and output staging tiles and the TMA store's reads add up to 4.6 MB per tile
// against 3.8 MB of port capacity in the ideal MMA time.
__host__ __device__ constexpr int pair_threads(int tw) { return (4 + 2 * tw) * 32; }
constexpr int kPairCtrlRegs = 96, kPairTeamRegs = 200;  // must fit what the CTA was launched with: 128 x 96 + 256 x 200 <= 384 x 168 (the pool is the CTA's own registers)

#ifdef LICOS_PAIR_PROBES  // development: per-role cycle counters (tools/probe_conv.py); build with LICOS_NVCC_EXTRA=-DLICOS_PAIR_PROBES
#define PP_T0(v) const long long v = clock64()
#define PP_ADD(acc, v) (acc) += clock64() - (v)
#else
#define PP_T0(v) (void)0
#define PP_ADD(acc, v) (void)0
#endif

// XC = 32-channel chunks PER THREAD the epilogue is unrolled for (TW = 4: 4 / 6 / 8 for N <= 128 / 192 / 256; TW = 8: half)
template <int EPI, bool OUT_NHWC, int XC, int TW>
__global__ void __launch_bounds__(pair_threads(TW), 1) conv_igemm_pair_kernel(const __grid_constant__ ConvParams p) {
    constexpr bool kGdn = (EPI == LICOS_EPI_GDN || EPI == LICOS_EPI_IGDN);
    constexpr int kNT = pair_threads(TW);
    constexpr uint32_t kTeamThreads = TW * 32;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], b_full[kMaxSB], b_empty[kMaxSB];
    __shared__ uint64_t acc_full[2], acc_empty[2], norm_full[2], stg_full[2], g_full;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[512];
    __shared__ __align__(16) float beta_s[256];
    __shared__ int16_t w_taps_s[kMaxPasses][kMaxSlabs * kMaxTaps];
    __shared__ int n_taps_s[kMaxPasses];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + (size_t)p.sa * p.a_slot_bytes;
    uint8_t* staging_all = b_ring + (size_t)p.sb * p.b_slot_bytes;          // one tile per epilogue team
    uint8_t* gamma_s = staging_all + (size_t)p.n_teams * p.staging_bytes;   // this CTA's HALF of gamma: (N / 64) atoms of [N / 2][64]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader_cta = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int half_n = p.N / 2;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < p.sb; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 2u * kTeamThreads * (uint32_t)p.jobs_per_pass);
            mbar_init(&norm_full[i], 1);
            mbar_init(&stg_full[i], 2);
        }
        mbar_init(&g_full, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < p.N * p.n_split; i += kNT) bias_s[i] = (p.bias && i < p.out_c) ? p.bias[i] : 0.f;
    if (kGdn)
        for (int i = threadIdx.x; i < p.N; i += kNT) beta_s[i] = p.beta[i];
    if (threadIdx.x < p.n_passes) {
        const Pass& ps = p.passes[threadIdx.x];
        int n = 0;
        for (int sl = 0; sl < ps.n_slabs; ++sl)
            for (int k = 0; k < ps.slabs[sl].n_taps; ++k) w_taps_s[threadIdx.x][n++] = ps.slabs[sl].taps[k].w_tap;
        n_taps_s[threadIdx.x] = n;
    }
    if (warp == 2) {
        tmem_alloc_pair(&tmem_base_smem, kTmemCols);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();  // both CTAs' barriers are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int n_items = p.total_tiles;  // pair work items (spatial tile pairs x channel splits)

    // GDN epilogues, TW = 4: the four control warps are exactly warpgroup 0 and need few registers; handing theirs to the two
    // epilogue teams (warpgroups 1 and 2) lets a team keep x AND two 32-column TMEM pieces in registers, so the next piece is
    // in flight while this one is worked on (at the 168 registers of a 384-thread CTA there was room for one).
#ifdef LICOS_NO_EPI_PIPE  // A/B knob (compile time)
    constexpr bool kPipe = false;
#else
    constexpr bool kPipe = kGdn && TW == 4 && XC <= 4;  // (N = 192 keeps 96 words of x: no room for a second piece)
#endif
    // (each setmaxnreg sits at the top of its role's branch: ptxas budgets the code that FOLLOWS it in that branch)
    if (warp < 4) {
    if (kPipe) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kPairCtrlRegs));
    if (warp == 0 && lane == 0) {
        // ===================== A producer (each CTA loads the slabs of its own tile) =====================
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.in_maps[i]);
        Ring ra;
        int n_loaded = 0;
        for (int w = pair; w < n_items; w += n_pairs) {
            const PairTile t = decode_pair_tile(p, w, rank);
            for (int pi = 0; pi < p.n_passes; ++pi) {
                const Pass& ps = p.passes[pi];
                for (int c = 0; c < p.cin_chunks; ++c) {
                    for (int s = 0; s < ps.n_slabs; ++s) {
                        const Slab& sl = ps.slabs[s];
                        mbar_wait(&a_empty[ra.slot], ra.phase ^ 1u);
                        if ((p.dbg_flags & 1) && n_loaded >= p.sa) {  // development: every ring slot loaded only once
                            if (leader_cta) mbar_arrive(&a_full[ra.slot]);
                            ra.advance(p.sa);
                            continue;
                        }
                        ++n_loaded;
                        if (leader_cta) mbar_arrive_expect_tx(&a_full[ra.slot], 2u * p.a_tx_bytes);
                        tma_load_4d_pair(a_ring + (size_t)ra.slot * p.a_slot_bytes, &p.in_maps[sl.in_map],
                                         mapa_shared(smem_u32(&a_full[ra.slot]), 0), c * kKChunk, t.gw0 + sl.dw, t.gh0 - 1, t.b);
                        ra.advance(p.sa);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== B producer: this CTA's half (N / 2 rows) of every weight tile =====================
        if (lane == 0) {
            tma_prefetch_desc(&p.w_map);
            if (kGdn) {  // this CTA's half of gamma stays resident for the whole kernel
                tma_prefetch_desc(&p.g_map);
                if (leader_cta) mbar_arrive_expect_tx(&g_full, 2u * p.gamma_bytes);
                const uint32_t gbar = mapa_shared(smem_u32(&g_full), 0);
                for (int gc = 0; gc < p.N / kKChunk; ++gc)
                    tma_load_2d_pair(gamma_s + (size_t)gc * half_n * 128, &p.g_map, gbar, gc * kKChunk, (int)rank * half_n);
            }
        }
        __syncwarp();
        uint32_t slot = 0, phase = 0;
        int n_loaded = 0;
        for (int w = pair; w < n_items; w += n_pairs) {
            const int ns_row = (w % p.n_split) * p.N + (int)rank * half_n;
            for (int pi = 0; pi < p.n_passes; ++pi) {
                const int nt = n_taps_s[pi];
                for (int c = 0; c < p.cin_chunks; ++c) {
                    for (int i = 0; i < nt; ++i) {
                        const int row = (int)w_taps_s[pi][i] * p.w_rows_per_tap + ns_row;
                        mbar_wait_warp(&b_empty[slot], phase ^ 1u);
                        if (elect_one()) {
                            if ((p.dbg_flags & 2) && n_loaded >= p.sb) {
                                if (leader_cta) mbar_arrive(&b_full[slot]);
                            } else {
                                if (leader_cta) mbar_arrive_expect_tx(&b_full[slot], 2u * p.b_slot_bytes);
                                tma_load_2d_pair(b_ring + (size_t)slot * p.b_slot_bytes, &p.w_map,
                                                 mapa_shared(smem_u32(&b_full[slot]), 0), c * kKChunk, row);
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
        // ===================== MMA issuer: the leader CTA issues for the pair =====================
        if (leader_cta) {
            const uint32_t idesc = umma_idesc_bf16(256, p.N);
            const uint64_t desc_hi = umma_desc_sw128(0);
            const uint32_t a_ring_addr = smem_u32(a_ring) >> 4, b_ring_addr = smem_u32(b_ring) >> 4;
            const uint32_t a_slot16 = p.a_slot_bytes >> 4, b_slot16 = p.b_slot_bytes >> 4;
            const uint32_t n_acc = p.n_acc, N = p.N, sa = p.sa, sb = p.sb;
            const uint32_t acc_step16 = p.acc_step16, pitch16 = p.a_pitch16;
            // the A operand's stride between 8-row groups is the geometry's (1 KB inside a 16-pixel slab row, or the whole row
            // pitch of a wide slab); weights keep the standard K-major tile
            const uint64_t desc_hi_a = ((uint64_t)1 << 16) | ((uint64_t)p.a_sbo16 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0, pit = 0;
#ifdef LICOS_PAIR_PROBES
            long long w_a = 0, w_b = 0, w_acc = 0, n_tiles = 0;
            const long long t_begin = clock64();
#endif
            // Look-ahead barrier checks.  A completed mbarrier wait costs this thread ~150 cycles, and tcgen05.mma issue does
            // not run far enough ahead of the pipe to hide them: one wait per tap and one per slab idled the tensor pipe 40 %
            // of the time (tools/probe_sweep.sh: the same with every load switched off).  So the NEXT barrier is probed with a
            // non-blocking, RELAXED test_wait BEFORE the current tap's MMAs are issued (an acquire there would serialise the
            // MMAs behind its latency) and its predicate is only read AFTER them: the latency runs under the 8 MMAs.  The
            // acquire is a fence on the consuming side; the blocking wait remains as the (rare) slow path.
            asm volatile(".reg .pred licos_nb, licos_na;");
            const uint32_t a_full0 = smem_u32(&a_full[0]), b_full0 = smem_u32(&b_full[0]);
            uint32_t have_a = 0, have_b = 0;  // 1 = the look-ahead probe of the current slot already saw its phase complete

            auto wait_slow = [&](uint32_t bar, uint32_t parity) {  // the look-ahead probe missed: a blocking (bounded) wait
                if (__all_sync(0xffffffffu, mbar_try_wait_addr(bar, parity) ? 1u : 0u)) return;
                const long long t0 = clock64();
                while (!__all_sync(0xffffffffu, mbar_try_wait_addr(bar, parity) ? 1u : 0u)) {
                    if (clock64() - t0 > 8000000000LL) __trap();  // a pipeline bug becomes a CUDA error, not a hung GPU
                }
            };

            for (int w = pair; w < n_items; w += n_pairs) {
#ifdef LICOS_PAIR_PROBES
                ++n_tiles;
#endif
                for (int pi = 0; pi < p.n_passes; ++pi, ++pit) {
                    const Pass& ps = p.passes[pi];
                    const uint32_t buf = pit % (uint32_t)p.n_buf;
                    {
                        PP_T0(_t);
                        mbar_wait_warp_cluster(&acc_empty[buf], ((pit / (uint32_t)p.n_buf) & 1u) ^ 1u);
                        PP_ADD(w_acc, _t);
                    }
                    tc_fence_after();
                    const uint32_t tmem_set = tmem_base + buf * (uint32_t)(ps.n_groups * (int)n_acc) * N;
                    const int n_slabs = ps.n_slabs;
                    uint32_t accumulate = 0;
                    for (int c = 0; c < p.cin_chunks; ++c) {
                        for (int s = 0; s < n_slabs; ++s) {
                            unsigned long long taps = p.tap_list[pi][s];  // [0:4) count, then (row_off : 2, col_off : 2) per tap
                            const int nt = (int)(taps & 15u);
                            taps >>= 4;
                            const uint32_t a_slab = a_ring_addr + a_slot * a_slot16;
                            if (!__all_sync(0xffffffffu, have_a)) {
                                PP_T0(_t);
                                wait_slow(a_full0 + a_slot * 8u, a_phase);
                                PP_ADD(w_a, _t);
                            }
                            const uint32_t na_slot = (a_slot + 1 == sa) ? 0u : a_slot + 1;
                            const uint32_t na_phase = a_phase ^ ((na_slot == 0u) ? 1u : 0u);
                            for (int k = 0; k < nt; ++k) {
                                if (!__all_sync(0xffffffffu, have_b)) {
                                    PP_T0(_t);
                                    wait_slow(b_full0 + b_slot * 8u, b_phase);
                                    PP_ADD(w_b, _t);
                                }
                                asm volatile("fence.acq_rel.cta;" ::: "memory");  // pairs with the relaxed look-ahead probes
                                tc_fence_after();
                                const uint64_t bd = desc_hi | (uint64_t)(b_ring_addr + b_slot * b_slot16);
                                const uint64_t ad = desc_hi_a | (uint64_t)(a_slab + ((uint32_t)taps & 3u) * pitch16 + (((uint32_t)taps >> 2) & 3u) * 8u);
                                taps >>= 4;
                                const uint32_t nb_slot = (b_slot + 1 == sb) ? 0u : b_slot + 1;
                                const uint32_t nb_phase = b_phase ^ ((nb_slot == 0u) ? 1u : 0u);
                                // probe the next tap's weights and, on a slab's last tap, the next slab -- the predicates are read
                                // after the MMAs below
                                asm volatile("mbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 licos_nb, [%0], %1;" ::"r"(b_full0 + nb_slot * 8u),
                                             "r"(nb_phase)
                                             : "memory");
                                if (k == nt - 1)
                                    asm volatile("mbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 licos_na, [%0], %1;" ::"r"(a_full0 + na_slot * 8u),
                                                 "r"(na_phase)
                                                 : "memory");
                                if (elect_one()) {
                                    if (n_acc == 2) {
                                        umma_bf16_pair(tmem_set, ad, bd, idesc, accumulate);
                                        umma_bf16_pair(tmem_set + N, ad + acc_step16, bd, idesc, accumulate);
#pragma unroll
                                        for (uint32_t ks = 1; ks < 4; ++ks) {
                                            umma_bf16_pair(tmem_set, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                                            umma_bf16_pair(tmem_set + N, ad + acc_step16 + 2 * ks, bd + 2 * ks, idesc, 1u);
                                        }
                                    } else {
                                        umma_bf16_pair(tmem_set, ad, bd, idesc, accumulate);
#pragma unroll
                                        for (uint32_t ks = 1; ks < 4; ++ks) umma_bf16_pair(tmem_set, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                                    }
                                    umma_commit_pair(&b_empty[b_slot]);
                                }
                                __syncwarp();
                                asm volatile("selp.b32 %0, 1, 0, licos_nb;" : "=r"(have_b));
                                accumulate = 1;
                                b_slot = nb_slot;
                                b_phase = nb_phase;
                            }
                            if (elect_one()) umma_commit_pair(&a_empty[a_slot]);
                            __syncwarp();
                            asm volatile("selp.b32 %0, 1, 0, licos_na;" : "=r"(have_a));
                            a_slot = na_slot;
                            a_phase = na_phase;
                        }
                    }
                    if (elect_one()) umma_commit_pair(&acc_full[buf]);
                    __syncwarp();
                }
            }
#ifdef LICOS_PAIR_PROBES
            if (p.dbg && lane == 0) {
                unsigned long long* d = p.dbg + blockIdx.x * 16;
                d[DBG_MMA_A] = w_a; d[DBG_MMA_B] = w_b; d[DBG_MMA_ACC] = w_acc; d[DBG_MMA_X2] = 0;
                d[DBG_MMA_TOTAL] = clock64() - t_begin; d[DBG_TILES] = n_tiles;
            }
#endif
        }
    }
    } else {
        if (kPipe) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kPairTeamRegs));
        // ===================== epilogue: two teams of TW warps per CTA, alternate accumulators =====================
        const int team = (warp - 4) / TW;
        const int half = TW == 8 ? (((warp - 4) >> 2) & 1) : 0;  // TW = 8: this warp takes chunks cc with (cc & 1) == half
        constexpr int kStride = TW == 8 ? 2 : 1;
        const int et = (warp & 3) * 32 + lane;   // == TMEM lane == row of the 128-row sub-tile
        const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
        const bool leader = et == 0 && half == 0;
        const bool first_warp = ((warp - 4) % TW) == 0;
        const int th = p.wide ? et >> 3 : et >> 4, tw = p.wide ? (et & 7) : (et & 15);  // position of this TMEM lane in its accumulator block
        uint8_t* staging = staging_all + (size_t)team * p.staging_bytes;
        uint32_t pit = 0, nit = 0, job = 0;
        const int n32 = p.N / 32;
        const uint32_t acc_empty_leader[2] = {mapa_shared(smem_u32(&acc_empty[0]), 0), mapa_shared(smem_u32(&acc_empty[1]), 0)};
        const uint32_t stg_full_leader = mapa_shared(smem_u32(&stg_full[team]), 0);
        const uint32_t idesc = umma_idesc_bf16(256, p.N);
        const uint32_t staging16 = smem_u32(staging) >> 4, gamma16 = smem_u32(gamma_s) >> 4;
        bool gamma_ready = false;
#ifdef LICOS_PAIR_PROBES
        const bool eprobe = p.dbg && leader && team == 0 && leader_cta;
        long long e_acc = 0, e_s1 = 0, e_norm = 0, e_s2 = 0, e_store = 0;
        const long long e_begin = clock64();
#endif
        for (int w = pair; w < n_items; w += n_pairs) {
            const PairTile t = decode_pair_tile(p, w, rank);
            const float* bias_t = bias_s + t.ns * p.N;
            for (int pi = 0; pi < p.n_passes; ++pi, ++pit) {
                const Pass& ps = p.passes[pi];
                const uint32_t buf = pit % (uint32_t)p.n_buf;
                const uint32_t tmem_set = tmem_base + lane_sel + buf * (uint32_t)(ps.n_groups * p.n_acc) * p.N;
                bool waited = false;
                for (int a = 0; a < p.n_acc; ++a) {
                    if (((job++) & (uint32_t)(p.n_teams - 1)) != (uint32_t)team) continue;
                    if (!waited) {
                        PP_T0(_t);
                        mbar_wait(&acc_full[buf], (pit / (uint32_t)p.n_buf) & 1u);
                        PP_ADD(e_acc, _t);
                        tc_fence_after();
                        waited = true;
                    }
                    const uint32_t acc = (uint32_t)a;
                    const uint32_t t_acc = tmem_set + acc * p.N;

                    if (OUT_NHWC && !kGdn) {
                        // this team's previous TMA store must have finished reading `staging` before it is rewritten
                        PP_T0(_t);
                        if (leader) tma_store_wait_read();
                        named_bar_sync(1 + team, kTeamThreads);
                        PP_ADD(e_store, _t);
                    }
                    uint32_t xs[kGdn ? XC * 16 : 1];  // v = acc + bias kept as packed bf16 pairs
                    float vp[kPipe ? 2 : 1][32];
                    if (kGdn) {
                        // stage 1: v^2 (bf16) -> staging = this CTA's 128 rows of the gamma GEMM's A operand; the pair-wide
                        // GEMM then overwrites both CTAs' accumulators IN PLACE with the norm
                        PP_T0(_s1);
                        if (kPipe) tmem_ld32(t_acc + half * 32, vp[0]);
#pragma unroll
                        for (int i = 0; i < XC; ++i) {
                            const int cc = kStride * i + half;
                            if (cc < n32) {
                                if (!kPipe) tmem_ld32(t_acc + cc * 32, vp[0]);
                                tmem_ld_wait();
                                if (kPipe && cc + kStride < n32) tmem_ld32(t_acc + (cc + kStride) * 32, vp[kPipe ? ((i + 1) & 1) : 0]);
                                uint32_t sq[16];
                                gdn_stage1_32<true>(vp[kPipe ? (i & 1) : 0], bias_t + cc * 32, xs + i * 16, sq);
                                if (i == 0) {
                                    // (GDN: the wait for the previous TMA store's read of `staging` comes only here, after
                                    // the first piece has been loaded and squared: it overlaps that work)
                                    if (OUT_NHWC && leader) tma_store_wait_read();
                                    named_bar_sync(1 + team, kTeamThreads);
                                }
                                store_row32(staging, et, cc, sq);
                            }
                        }
                        fence_proxy_async();
                        tc_fence_before();
                        named_bar_sync(1 + team, kTeamThreads);
                        if (first_warp) {
                            // "staged" across the pair; once both CTAs are, the leader's warp issues the pair-wide gamma GEMM.
                            // (Issuing it from the main MMA thread instead -- it owns the pipe's queue -- was measured and was
                            // no faster: in pair mode the kernel is bound by the 128 B/clk shared-memory port, not by issue.)
#ifdef LICOS_STG_RELEASE  // A/B knob (compile time): the cluster-scope release arrive of the first version
                            if (elect_one()) mbar_arrive_cluster(stg_full_leader);
#else
                            // The staged tile never leaves this SM: cta_group::2 makes THIS SM's tensor core read it, and the
                            // stores were fenced for the async proxy and ordered by the team barrier above.  What crosses to
                            // the leader CTA is only the signal, so it needs no cluster-scope release (~1.5 k cycles).
                            if (elect_one()) mbar_arrive_cluster_relaxed(stg_full_leader);
#endif
                            __syncwarp();
                            if (leader_cta) {
                                if (!gamma_ready) { mbar_wait(&g_full, 0); gamma_ready = true; }
                                mbar_wait_warp_cluster(&stg_full[team], nit & 1u);
                                tc_fence_after();
                                if (elect_one()) {
                                    issue_gamma_gemm_pair_n(tmem_set - lane_sel + acc * p.N, staging16, gamma16, (uint32_t)p.N, idesc);
                                    umma_commit_pair(&norm_full[team]);
                                }
                                __syncwarp();
                            }
                        }
                        PP_ADD(e_s1, _s1);
                        PP_T0(_n);
                        mbar_wait(&norm_full[team], nit & 1u);
                        PP_ADD(e_norm, _n);
                        tc_fence_after();
                        ++nit;
                        if (kPipe) tmem_ld32(t_acc + half * 32, vp[0]);
                    }

                    // stage 2: activation, then write out
                    PP_T0(_s2);
                    const int bh0 = t.gh0 + (p.wide ? 0 : a * kAccRows), bw0 = t.gw0 + (p.wide ? 8 * a : 0);  // the block's origin
                    const int gh = bh0 + th, gw = bw0 + tw;
                    const int oh = gh * p.out_s + ps.dy[0], ow = gw * p.out_s + ps.dx[0];
                    const bool in_range = t.live && gh < p.grid_h && gw < p.grid_w && oh < p.out_h && ow < p.out_w;
                    const size_t cs = (size_t)p.out_h * p.out_w;
                    float* o = OUT_NHWC ? nullptr
                                        : p.out_f32 + ((size_t)t.b * p.out_c * p.out_h + oh) * p.out_w + ow +
                                              (size_t)(t.ns * p.N) * cs;
                    const int c_left = p.out_c - t.ns * p.N;  // valid channels from this split's base
                    __nv_bfloat16* pre_px = (kGdn && p.pre_out && in_range)
                                                ? p.pre_out + (((size_t)t.b * p.out_h + oh) * p.out_w + ow) * p.out_c + t.ns * p.N
                                                : nullptr;
#pragma unroll
                    for (int i = 0; i < XC; ++i) {
                        const int cc = kStride * i + half;
                        if (cc < n32) {
                            float vs[kPipe ? 1 : 32];
                            if constexpr (kPipe) {
                                tmem_ld_wait();
                                if (cc + kStride < n32) {
                                    tmem_ld32(t_acc + (cc + kStride) * 32, vp[(i + 1) & 1]);
                                } else {
#ifndef LICOS_LATE_ACC_RELEASE  // (A/B knob, compile time)
                                    // the last piece is in registers: this thread is done with the accumulator -- hand it back
                                    // to the pair's MMA issuer now, one piece of arithmetic and stores earlier
                                    tc_fence_before();
                                    mbar_arrive_cluster_relaxed(acc_empty_leader[buf]);
#endif
                                }
                            } else {
                                tmem_ld32(t_acc + cc * 32, vs);  // GDN: the norm; otherwise the accumulator
                                tmem_ld_wait();
                            }
                            float (&v)[32] = *reinterpret_cast<float (*)[32]>(kPipe ? &vp[kPipe ? (i & 1) : 0][0] : &vs[0]);
                            if (kGdn) {
                                if (pre_px && cc * 32 < c_left) {
#pragma unroll
                                    for (int q = 0; q < 4; ++q)
                                        reinterpret_cast<uint4*>(pre_px + cc * 32)[q] =
                                            make_uint4(xs[i * 16 + 4 * q], xs[i * 16 + 4 * q + 1], xs[i * 16 + 4 * q + 2], xs[i * 16 + 4 * q + 3]);
                                }
                                uint32_t out[16];
                                gdn_stage2_32<EPI == LICOS_EPI_IGDN>(v, beta_s + cc * 32, xs + i * 16, out);
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
                                        for (int ii = 0; ii < 4; ++ii) v[4 * q + ii] = fmaxf(v[4 * q + ii], 0.f);
                                    }
                                }
                                if (OUT_NHWC) {
                                    uint32_t out[16];
#pragma unroll
                                    for (int ii = 0; ii < 16; ++ii) out[ii] = pack_bf16x2(v[2 * ii], v[2 * ii + 1]);
                                    store_row32(staging, et, cc, out);
                                } else if (in_range) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (cc * 32 + j < c_left) o[(size_t)(cc * 32 + j) * cs] = v[j];
                                }
                            }
                        }
                    }
                    if (!OUT_NHWC && !kGdn && (p.N & 16) && half == (TW == 8 ? (n32 & 1) : 0)) {  // trailing 16 columns (N % 32 == 16)
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
                    // this accumulator has been read: hand it back to the pair's MMA issuer
#ifndef LICOS_LATE_ACC_RELEASE
                    if (!kPipe)
#endif
                    {
                        tc_fence_before();
                        mbar_arrive_cluster_relaxed(acc_empty_leader[buf]);  // orders TMEM reads (the tcgen05 fence), no memory
                    }
                    PP_ADD(e_s2, _s2);
                    PP_T0(_st);
                    if (OUT_NHWC) {
                        fence_proxy_async();
                        named_bar_sync(1 + team, kTeamThreads);
                        if (leader) {
                            for (int at = 0; at < p.N / kKChunk; ++at) {
                                tma_store_4d(&p.out_maps[ps.out_map[0]], staging + (size_t)at * (128 * 128),
                                             at * kKChunk, bw0, bh0, t.b);
                            }
                            tma_store_commit();
                        }
                    }
                    PP_ADD(e_store, _st);
                }
            }
        }
        if (leader) tma_store_wait_all();
#ifdef LICOS_PAIR_PROBES
        if (eprobe) {
            unsigned long long* d = p.dbg + blockIdx.x * 16;
            d[DBG_EPI_ACC] = e_acc; d[DBG_EPI_S1] = e_s1; d[DBG_EPI_NORM] = e_norm; d[DBG_EPI_S2] = e_s2;
            d[DBG_EPI_STORE] = e_store; d[DBG_EPI_TOTAL] = clock64() - e_begin;
        }
#endif
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync();  // neither CTA leaves (or frees TMEM) while the other may still reach into it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

}  // namespace licos
