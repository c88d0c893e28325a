// Which part of the conv mainloop's per-tap protocol stalls the tensor pipe?  M=128 N=128 K=16, 8 MMAs per "tap".
#include <cstdio>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

// mode bits: 1 = commit per tap, 2 = try_wait (already complete) + fence per tap, 4 = alternate 2 accumulators,
//            8 = clock64 probe pair per tap, 16 = 90-cycle mbarrier wait replaced by plain fence only
__global__ void __launch_bounds__(128, 1) k(int mode, int taps, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t done, ready[4], freed[4];
    __shared__ uint32_t tb;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        mbar_init(&done, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&ready[i], 1); mbar_init(&freed[i], 1); }
        mbar_fence_init();
        for (int i = 0; i < 4; ++i) mbar_arrive(&ready[i]);  // phase 0 complete: wait(parity 0) succeeds at once
    }
    if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint64_t hi = umma_desc_sw128(0);
        const uint32_t a16 = base >> 4, b16 = (base + 110592) >> 4;
        long long sink = 0;
        const long long t0 = clock64();
        for (int t = 0; t < taps; ++t) {
            const int s = t & 3;
            if (mode & 2) { mbar_wait(&ready[s], 0); tc_fence_after(); }
            if (mode & 16) tc_fence_after();
            long long c0 = 0;
            if (mode & 8) c0 = clock64();
            const uint64_t bd = hi | (uint64_t)(b16 + s * 1024);
            for (int a = 0; a < 2; ++a) {
                const uint64_t ad = hi | (uint64_t)(a16 + (t % 3) * 2304 + a * 1024);
                const uint32_t d = tmem + ((mode & 4) ? a * 128 : 0);
                umma_bf16(d, ad, bd, idesc, 1u);
                umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
                umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
            }
            if (mode & 1) umma_commit(&freed[s]);
            if (mode & 8) sink += clock64() - c0;
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        out[blockIdx.x] = clock64() - t0 + (sink & 1);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int taps = 1000;
    for (int mode : {0, 1, 2, 3, 4, 7, 15, 16, 17}) {
        k<<<148, 128, 200 * 1024>>>(mode, taps, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("mode=%2d (commit=%d wait+fence=%d alt_acc=%d probe=%d fence_only=%d): %.1f cycles/MMA, %.0f cycles/tap [%s]\n", mode, mode & 1,
               (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, (mode >> 4) & 1, avg / (taps * 8.0), avg / taps, cudaGetErrorString(e));
    }
    return 0;
}
