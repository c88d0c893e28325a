// Pipe behaviour vs operand order inside a tap of 8 MMAs (warp-uniform lean issue loop, no waits):
//  M0 (d0,k0)(d1,k0)(d0,k1)(d1,k1)...   two accumulators interleaved, each B slice used twice in a row
//  M1 (d0,k0..k3)(d1,k0..k3)            4 + 4
//  M2 burst-like: accumulator alternates, A atom = i & 3, k = i & 3 for both operands
//  M3 (d0,k0)(d1,k1)(d0,k2)(d1,k3)(d0,k1)(d1,k2)(d0,k3)(d1,k0): no back-to-back reuse of a B slice or an A slice
//  M4 one accumulator only, k cycles
#include <cstdio>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int taps, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t done;
    __shared__ uint32_t tb;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    if (threadIdx.x < 32) {
        const uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint64_t hi = umma_desc_sw128(0);
        const uint32_t a16 = base >> 4, b16 = (base + 81920) >> 4;
        const long long t0 = clock64();
        uint32_t slot = 0, arow = 0;
        for (int t = 0; t < taps; ++t) {
            const uint64_t bd = hi | (uint64_t)(b16 + slot * 1024);
            const uint64_t ad = hi | (uint64_t)(a16 + arow * 128);
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int acc, atom, ka, kb;
                    if (MODE == 0) { acc = i & 1; atom = i & 1; ka = kb = i >> 1; }
                    else if (MODE == 1) { acc = i >> 2; atom = i >> 2; ka = kb = i & 3; }
                    else if (MODE == 2) { acc = i & 1; atom = i & 3; ka = kb = i & 3; }
                    else if (MODE == 3) { acc = i & 1; atom = i & 1; ka = kb = (i + (i >> 2)) & 3; }
                    else { acc = 0; atom = 0; ka = kb = i & 3; }
                    umma_bf16(tmem + acc * 128, ad + atom * 1024 + 2 * ka, bd + 2 * kb, idesc, 1u);
                }
            }
            __syncwarp();
            slot = (slot == 4) ? 0 : slot + 1;
            arow = (arow == 2) ? 0 : arow + 1;
        }
        if (elect_one()) umma_commit(&done);
        __syncwarp();
        mbar_wait(&done, 0);
        if (threadIdx.x == 0) out[0] = (clock64() - t0) / taps;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE> void run(long long* d) {
    long long h[1];
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<MODE><<<1, 128, 200 * 1024>>>(2000, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("order M%d: %lld cycles/tap = %.1f/MMA [%s]\n", MODE, h[0], h[0] / 8.0, cudaGetErrorString(e));
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    run<0>(d); run<1>(d); run<2>(d); run<3>(d); run<4>(d);
    return 0;
}
