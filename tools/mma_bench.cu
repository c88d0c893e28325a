// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, cta_group::1, SS mode) as a function of N,
// operands resident in shared memory, no loads.  Decides how the conv engine should shape its MMAs.
#include <cstdio>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

__global__ void __launch_bounds__(128, 1) k(int N, int iters, int n_acc, int b_stride16, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tb;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint64_t hi = umma_desc_sw128(0);
        const uint32_t a16 = base >> 4, b16 = (base + 65536) >> 4;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t acc = (uint32_t)(i % n_acc);
            const uint64_t ad = hi | (uint64_t)(a16 + (i & 3) * 1024);            // 4 different 16 KB A tiles
            const uint64_t bd = hi | (uint64_t)(b16 + (i & 3) * b_stride16);
            umma_bf16(tmem + acc * N, ad, bd, idesc, 1u);
            umma_bf16(tmem + acc * N, ad + 2, bd + 2, idesc, 1u);
            umma_bf16(tmem + acc * N, ad + 4, bd + 4, idesc, 1u);
            umma_bf16(tmem + acc * N, ad + 6, bd + 6, idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    for (int grid : {1, 148}) for (int n_acc : {1, 2}) for (int N : {16, 64, 128, 256}) {
        if (n_acc * N > 512) continue;
        k<<<grid, 128, 200 * 1024>>>(N, iters, n_acc, (N * 128) >> 4, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
        printf("grid=%3d n_acc=%d N=%3d: %.1f cycles/MMA (ideal %.1f) smem bytes/MMA=%d -> %.1f B/cyc  [%s]\n", grid, n_acc, N,
               avg / (iters * 4.0), 128.0 * N / 256.0, 4096 + N * 32, (4096 + N * 32) / (avg / (iters * 4.0)), cudaGetErrorString(e));
    }
    return 0;
}
