// How deep is the tcgen05.mma issue queue, and what do the per-tap protocol pieces cost the issuing thread?
//  (a) burst of K MMAs (M=128, N=128, K=16) on an idle pipe: cycles until the LAST ISSUE returns vs until completion
//  (b) cost, for the issuing thread, of: try_wait on a completed mbarrier, tcgen05.commit, tcgen05.fence::after
#include <cstdio>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

__global__ void __launch_bounds__(128, 1) k(int burst, int reps, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t done, ready;
    __shared__ uint32_t tb;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_init(&ready, 1); mbar_fence_init(); mbar_arrive(&ready); }
    if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint64_t hi = umma_desc_sw128(0);
        const uint32_t a16 = base >> 4, b16 = (base + 65536) >> 4;
        long long t_issue = 0, t_done = 0, t_wait = 0, t_commit = 0, t_fence = 0;
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            for (int i = 0; i < burst; ++i)
                umma_bf16(tmem + (i & 1) * 128, hi | (uint64_t)(a16 + (i & 3) * 1024 + (i & 3) * 2), hi | (uint64_t)(b16 + (i & 3) * 2), idesc, 1u);
            const long long t1 = clock64();
            umma_commit(&done);
            const long long t2 = clock64();
            mbar_wait(&done, phase);
            phase ^= 1u;
            const long long t3 = clock64();
            t_issue += t1 - t0; t_commit += t2 - t1; t_done += t3 - t0;
            const long long t4 = clock64();
            mbar_wait(&ready, 0);  // already complete
            const long long t5 = clock64();
            tc_fence_after();
            const long long t6 = clock64();
            t_wait += t5 - t4; t_fence += t6 - t5;
        }
        out[0] = t_issue / reps; out[1] = t_done / reps; out[2] = t_commit / reps; out[3] = t_wait / reps; out[4] = t_fence / reps;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// steady state: per "tap" = [optional wait on a completed barrier] + G MMAs + [optional commit]; taps back to back
__global__ void __launch_bounds__(128, 1) k2(int G, int taps, int do_wait, int do_commit, int lookahead, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t done, ready[8], freed[8];
    __shared__ uint32_t tb;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        mbar_init(&done, 1);
        for (int i = 0; i < 8; ++i) { mbar_init(&ready[i], 1); mbar_init(&freed[i], 1); }
        mbar_fence_init();
        for (int i = 0; i < 8; ++i) mbar_arrive(&ready[i]);
    }
    if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint64_t hi = umma_desc_sw128(0);
        const uint32_t a16 = base >> 4, b16 = (base + 65536) >> 4;
        const long long t0 = clock64();
        if (lookahead && do_wait) { mbar_wait(&ready[0], 0); }
        for (int t = 0; t < taps; ++t) {
            const int s = t & 7;
            if (do_wait && !lookahead) { mbar_wait(&ready[s], 0); tc_fence_after(); }
            const uint64_t bd = hi | (uint64_t)(b16 + (s & 3) * 1024);
            for (int g = 0; g < G; ++g) {
                const uint64_t ad = hi | (uint64_t)(a16 + ((g >> 2) & 1) * 1024 + (g & 3) * 2);
                umma_bf16(tmem + ((g >> 2) & 1) * 128, ad, bd + (g & 3) * 2, idesc, 1u);
            }
            if (do_commit) umma_commit(&freed[s]);
            if (do_wait && lookahead) { mbar_wait(&ready[(t + 1) & 7], 0); tc_fence_after(); }  // next tap's wait AFTER the issue
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        out[0] = (clock64() - t0) / taps;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    long long h[8];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int burst : {1, 2, 4, 8, 16, 32, 64}) {
        k<<<1, 128, 200 * 1024>>>(burst, 50, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
        printf("burst=%2d: issue %lld cyc (%.1f/MMA), issue->complete %lld cyc (%.1f/MMA), commit %lld, wait(complete) %lld, fence %lld [%s]\n",
               burst, h[0], (double)h[0] / burst, h[1], (double)h[1] / burst, h[2], h[3], h[4], cudaGetErrorString(e));
    }
    for (int G : {4, 8, 16})
        for (int mode = 0; mode < 5; ++mode) {
            const int do_wait = mode >= 1 && mode != 2, do_commit = mode >= 2, look = mode == 4;
            k2<<<1, 128, 200 * 1024>>>(G, 2000, do_wait, do_commit, look, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            printf("steady G=%2d wait=%d commit=%d lookahead=%d: %lld cycles/tap = %.1f cycles/MMA [%s]\n", G, do_wait, do_commit, look, h[0],
                   (double)h[0] / G, cudaGetErrorString(e));
        }
    return 0;
}
