// What costs the conv engine its tensor-pipe cycles?  The engine's issue loop on EVERY SM at once (the earlier probes ran one
// CTA), with the other actors of the real kernel switched on one at a time (bit flags):
//   1   an (already completed) mbarrier wait + tcgen05.fence before every tap of 8 MMAs      -> pure issue-side latency
//   2   a tcgen05.commit to an mbarrier after every tap                                       -> as the real loop
//   4   four "epilogue" warps streaming tcgen05.ld of a second TMEM region + 16-byte st.shared -> TMEM / smem port contention
//   8   a producer thread streaming cp.async.bulk (TMA engine) global -> smem, 18 KB per tap  -> smem write traffic of the slabs
//   16  CTA pairs: tcgen05.mma.cta_group::2, M = 256 (128 rows per SM), each SM supplies half of B (64 of the N = 128 rows)
//   32  N = 256 instead of 128 (per-MMA operand bytes per cycle halve for A)
// Prints cycles per MMA (SM clock) and the wall-clock rate, so clock throttling under full-chip tensor load is visible too.
//   usage: mma_bench7 [taps]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(192, 1) bench(int taps, int flags, const uint8_t* __restrict__ gsrc, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t done, tapbar, ready, ring_full[4];
    __shared__ uint32_t tb;
    __shared__ volatile int stop;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* smem = raw + (base - smem_u32(raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) {
        mbar_init(&done, 1); mbar_init(&tapbar, 1); mbar_init(&ready, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&ring_full[i], 1);
        mbar_fence_init();
        stop = 0;
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(&tb, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tb;
    const int N = (flags & 32) ? 256 : 128;
    if (threadIdx.x == 0) mbar_arrive(&ready);  // a completed phase to wait on (flag 1)

    if (warp == 0) {
        if (!PAIR || rank == 0) {
            const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, N);
            const uint64_t hi = umma_desc_sw128(0);
            const uint32_t a16 = base >> 4, b16 = (base + 73728) >> 4;  // A: 2 x 36 KB slabs, B: 4 slots of N x 128 B (N/2 rows per SM in a pair)
            const uint32_t b_slot16 = (uint32_t)(PAIR ? N / 2 : N) * 8u;
            const long long t0 = clock64();
            uint32_t slot = 0, arow = 0, tphase = 0;
            for (int t = 0; t < taps; ++t) {
                if (flags & 1) {
                    mbar_wait_warp(&ready, 0);
                    tc_fence_after();
                }
                const uint64_t bd = hi | (uint64_t)(b16 + slot * b_slot16);
                const uint64_t ad = hi | (uint64_t)(a16 + arow * 128);
                if (elect_one()) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int acc = i & 1, ks = i >> 1;
                        if (PAIR) umma_bf16_2cta(tmem + acc * N, ad + acc * 1024 + 2 * ks, bd + 2 * ks, idesc, 1u);
                        else umma_bf16(tmem + acc * N, ad + acc * 1024 + 2 * ks, bd + 2 * ks, idesc, 1u);
                    }
                    if (flags & 2) { if (PAIR) umma_commit_2cta(&tapbar); else umma_commit(&tapbar); }
                }
                __syncwarp();
                if ((flags & 2) && (t & 3) == 3) {  // keep the barrier's phases bounded: drain every 4 taps' commits... cheaply
                    (void)tphase;
                }
                slot = (slot == 3) ? 0 : slot + 1;
                arow = (arow == 2) ? 0 : arow + 1;
            }
            if (elect_one()) { if (PAIR) umma_commit_2cta(&done); else umma_commit(&done); }
            __syncwarp();
            mbar_wait(&done, 0);
            if (lane == 0) { out[2 * blockIdx.x] = (clock64() - t0); stop = 1; }
        } else {
            mbar_wait(&done, 0);  // the leader's final commit is multicast to both CTAs
            if (lane == 0) { out[2 * blockIdx.x] = 0; stop = 1; }
        }
    } else if (warp == 1) {
        // flag 8: TMA-engine write traffic into a 4-slot ring (not read by anyone: pure port contention)
        if ((flags & 8) && lane == 0) {
            uint8_t* ring = smem + 147456;  // 4 x 18 KB
            uint32_t n = 0;
            while (!stop) {
                const uint32_t s = n & 3u;
                if (n >= 4) mbar_wait(&ring_full[s], ((n >> 2) - 1) & 1u);
                mbar_arrive_expect_tx(&ring_full[s], 18432);
                bulk_g2s(ring + s * 18432, gsrc + ((size_t)((blockIdx.x * 131 + n) % 2048)) * 18432, 18432, &ring_full[s]);
                ++n;
            }
            for (uint32_t k = (n >= 4 ? n - 4 : 0); k < n; ++k) mbar_wait(&ring_full[k & 3u], (k >> 2) & 1u);
        }
    } else {
        // flag 4: epilogue-like traffic: TMEM -> registers -> 16-byte shared stores
        if (flags & 4) {
            const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
            uint8_t* stg = smem + 114688 + (warp - 2) * 4096;
            float acc = 0.f;
            long long iters = 0;
            const long long e0 = clock64();
            while (!stop) {
                float v[32];
                tmem_ld32(tmem + lane_sel + 256 + ((warp & 1) * 32), v);
                tmem_ld_wait();
                if (!(flags & 64)) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        *reinterpret_cast<float4*>(stg + ((lane * 8 + q) & 255) * 16) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        acc += v[4 * q];
                    }
                } else {
                    acc += v[0] + v[31];
                }
                ++iters;
            }
            if (warp == 2 && lane == 0) out[2 * blockIdx.x + 1] = (clock64() - e0) / (iters ? iters : 1);  // cycles per tcgen05.ld.x32 round trip
            if (acc == 12345.f) out[1] = 1;
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else tmem_dealloc(tmem, 512);
    }
}

static void run(int flags, int taps, int grid, const uint8_t* src, long long* d) {
    const bool pair = flags & 16;
    const int smem = 225 * 1024;
    cudaMemset(d, 0, 16 * 512);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaError_t err;
    if (pair) {
        cudaFuncSetAttribute(bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid & ~1); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaEventRecord(e0);
        err = cudaLaunchKernelEx(&cfg, bench<true>, taps, flags, src, d);
        cudaEventRecord(e1);
    } else {
        cudaFuncSetAttribute(bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaEventRecord(e0);
        bench<false><<<grid, 192, smem>>>(taps, flags, src, d);
        err = cudaGetLastError();
        cudaEventRecord(e1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[2 * 160];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0, ldsum = 0; int n = 0, nld = 0;
    for (int i = 0; i < grid; ++i) if (h[2 * i] > 0) { sum += (double)h[2 * i]; ++n; }
    for (int i = 0; i < grid; ++i) if (h[2 * i + 1] > 1) { ldsum += (double)h[2 * i + 1]; ++nld; }
    const double cyc = n ? sum / n / taps / 8.0 : 0;
    const int sms = pair ? (grid & ~1) : grid;
    const int Nn = (flags & 32) ? 256 : 128;
    const double tflops = 2.0 * 128 * Nn * 16 * 8.0 * taps * sms / (ms * 1e-3) / 1e12;
    printf("flags %3d grid %3d: %.1f cycles/MMA (ideal %d)  %.3f ms  %.0f TFLOP/s  eff clock %.2f GHz  tmem_ld32 loop %.0f cycles [%s %s]\n",
           flags, grid, cyc, Nn / 2, ms, tflops, n ? (sum / n) / (ms * 1e-3) / 1e9 : 0.0, nld ? ldsum / nld : 0.0, cudaGetErrorString(err),
           cudaGetErrorString(e));
}

int main(int argc, char** argv) {
    const int taps = argc > 1 ? atoi(argv[1]) : 20000;
    long long* d; cudaMalloc(&d, 16 * 512);
    uint8_t* src; cudaMalloc(&src, (size_t)2048 * 18432 + 65536); cudaMemset(src, 0, (size_t)2048 * 18432);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int sets[] = {0, 3, 4, 4 + 64, 15, 16, 19, 16 + 4, 16 + 4 + 64, 31, 31 + 64, 32, 48, 63};
    printf("== one CTA\n");
    run(0, taps, 1, src, d); run(3, taps, 1, src, d); run(15, taps, 1, src, d);
    printf("== all %d SMs\n", sms);
    for (int f : sets) run(f, taps, sms, src, d);
    return 0;
}
