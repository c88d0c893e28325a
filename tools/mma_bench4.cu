// Where does the tensor pipe stall in a conv-style issue loop?  One thread, taps of 8 unrolled MMAs (two accumulators
// interleaved), the operand pattern of the conv engine.  Variants switch one thing at a time; the first taps also record
// a timestamp after every MMA issue (issue is blocking, so the deltas are the pipe's acceptance times).
//   bit 0: B tile changes per tap          bit 1: A row offset changes per tap     bit 2: commit per tap
//   bit 3: wait (completed barrier) + fence per tap        bit 4: accumulators interleaved (else 4 + 4)
//   bit 5: wait placed after the first MMA pair of the tap instead of before it (for the NEXT tap's barrier)
#include <cstdio>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

__global__ void __launch_bounds__(128, 1) k(int mode, int taps, long long* out, long long* stamps, int a_off, int b_off, int pace) {
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
        const uint32_t a16 = (base + a_off) >> 4, b16 = (base + b_off) >> 4;
        const bool b_var = mode & 1, a_var = mode & 2, do_commit = mode & 4, do_wait = mode & 8, inter = mode & 16, late = mode & 32;
        uint32_t dummy = taps;
        auto spin = [&]() { for (int z = 0; z < pace; ++z) dummy = dummy * 1664525u + 1013904223u; };
        const long long t0 = clock64();
        for (int t = 0; t < taps; ++t) {
            const int s = t % 5;
            if (do_wait && !late) { mbar_wait(&ready[s & 7], 0); tc_fence_after(); }
            const uint64_t bd = hi | (uint64_t)(b16 + (b_var ? s * 1024 : 0));
            const uint64_t ad = hi | (uint64_t)(a16 + (a_var ? (t % 3) * 128 : 0));   // row offset = multiples of 2 KB
            const uint32_t d0 = tmem, d1 = tmem + 128;
            long long c[9];
            const bool rec = t >= 8 && t < 12;
            if (rec) c[0] = clock64();
            if (inter) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_bf16(d0, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                    spin();
                    if (rec) c[2 * ks + 1] = clock64();
                    umma_bf16(d1, ad + 1024 + 2 * ks, bd + 2 * ks, idesc, 1u);
                    spin();
                    if (rec) c[2 * ks + 2] = clock64();
                    if (ks == 0 && do_wait && late) { mbar_wait(&ready[(s + 1) & 7], 0); tc_fence_after(); }
                }
            } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) { umma_bf16(d0, ad + 2 * ks, bd + 2 * ks, idesc, 1u); if (rec) c[ks + 1] = clock64(); }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) { umma_bf16(d1, ad + 1024 + 2 * ks, bd + 2 * ks, idesc, 1u); if (rec) c[ks + 5] = clock64(); }
            }
            if (do_commit) umma_commit(&freed[s & 7]);
            if (rec) for (int i = 0; i < 9; ++i) stamps[(t - 8) * 9 + i] = c[i] - t0;
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        out[0] = (clock64() - t0) / taps;
        out[1] = dummy;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long *d, *st; cudaMalloc(&d, 64); cudaMalloc(&st, 36 * 8);
    long long h[1], hs[36];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode : {0, 16, 17, 18, 19, 20, 24, 28, 31, 63, 1, 3, 15}) {
        k<<<1, 128, 200 * 1024>>>(mode, 2000, d, st, 0, 81920, 0);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hs, st, 36 * 8, cudaMemcpyDeviceToHost);
        printf("mode=%2d (bvar=%d avar=%d commit=%d wait=%d interleave=%d late=%d): %lld cycles/tap = %.1f/MMA [%s]\n   issue deltas:", mode, mode & 1,
               (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, (mode >> 4) & 1, (mode >> 5) & 1, h[0], h[0] / 8.0, cudaGetErrorString(e));
        for (int t = 0; t < 3; ++t) {
            printf(" |");
            for (int i = 1; i < 9; ++i) printf(" %lld", hs[t * 9 + i] - hs[t * 9 + i - 1]);
            printf(" (gap to next tap %lld)", hs[(t + 1) * 9] - hs[t * 9 + 8]);
        }
        printf("\n");
    }
    for (int a_off : {0, 16384, 36864})
        for (int b_off : {32768, 49152, 65536, 73728, 81920, 98304, 131072, 163840}) {
            k<<<1, 128, 200 * 1024>>>(16, 2000, d, st, a_off, b_off, 0);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            printf("a_off=%6d b_off=%6d: %lld cycles/tap = %.1f/MMA [%s]\n", a_off, b_off, h[0], h[0] / 8.0, cudaGetErrorString(e));
        }
    for (int mode : {16, 31})
    for (int pace : {0, 1, 2, 3, 4, 6, 8, 12, 16}) {
        k<<<1, 128, 200 * 1024>>>(mode, 2000, d, st, 0, 81920, pace);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hs, st, 36 * 8, cudaMemcpyDeviceToHost);
        printf("mode=%d pace=%2d: %lld cycles/tap = %.1f/MMA [%s]  deltas:", mode, pace, h[0], h[0] / 8.0, cudaGetErrorString(e));
        for (int i = 1; i < 9; ++i) printf(" %lld", hs[i] - hs[i - 1]);
        printf("\n");
    }
    return 0;
}
