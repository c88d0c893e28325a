// Issue-loop shape: the same conv-style tap loop (8 MMAs per tap, completed-barrier wait + commit per tap) issued
//  (a) from inside `if (threadIdx.x == 0)` (divergent code: operands in vector registers, R2UR + ELECT loops per MMA)
//  (b) warp-uniformly: all 32 lanes run the loop and wait, the MMAs sit under elect_one() (operands can stay uniform)
#include <cstdio>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

template <bool UNIFORM>
__global__ void __launch_bounds__(128, 1) k(int taps, int do_wait, int do_commit, long long* out) {
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
    if (UNIFORM ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
        const uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint64_t hi = umma_desc_sw128(0);
        const uint32_t a16 = base >> 4, b16 = (base + 81920) >> 4;
        const long long t0 = clock64();
        uint32_t slot = 0, arow = 0;
        for (int t = 0; t < taps; ++t) {
            if (do_wait) { mbar_wait(&ready[slot], 0); tc_fence_after(); }
            const uint64_t bd = hi | (uint64_t)(b16 + slot * 1024);
            const uint64_t ad = hi | (uint64_t)(a16 + arow * 128);
            const uint32_t d0 = tmem, d1 = tmem + 128;
            if (!UNIFORM || elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_bf16(d0, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                    umma_bf16(d1, ad + 1024 + 2 * ks, bd + 2 * ks, idesc, 1u);
                }
                if (do_commit) umma_commit(&freed[slot]);
            }
            if (UNIFORM) __syncwarp();
            slot = (slot == 4) ? 0 : slot + 1;
            arow = (arow == 2) ? 0 : arow + 1;
        }
        if (!UNIFORM || elect_one()) umma_commit(&done);
        if (UNIFORM) __syncwarp();
        mbar_wait(&done, 0);
        if (threadIdx.x == 0) out[0] = (clock64() - t0) / taps;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    long long h[1];
    cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int uni = 0; uni < 2; ++uni)
        for (int w = 0; w < 2; ++w)
            for (int c = 0; c < 2; ++c) {
                if (uni) k<true><<<1, 128, 200 * 1024>>>(2000, w, c, d); else k<false><<<1, 128, 200 * 1024>>>(2000, w, c, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
                printf("uniform=%d wait=%d commit=%d: %lld cycles/tap = %.1f/MMA [%s]\n", uni, w, c, h[0], h[0] / 8.0, cudaGetErrorString(e));
            }
    return 0;
}
