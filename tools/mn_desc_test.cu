// Experiment: MN-major SWIZZLE_128B UMMA operands read straight from TMA-written [pixel][64 channel] tiles, with the
// operand starting at an arbitrary 128-byte pixel row (what licos_conv_wgrad's shifted tap views rely on).
// A = [P pixels][128 ch] as two 64-channel chunks (LBO apart), B = [P pixels][64 ch];
// D[m][n] = sum_{k < 64} A[sa + k][m] * B[sb + k][n] must hold for every start row sa, sb.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

constexpr int P = 96;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                                            int sa, int sb, float* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar, done;
    __shared__ uint32_t tb;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* a_s = sm;                 // 2 chunks of P * 128 bytes (chunk pitch 12 KB... P*128 = 12288)
    uint8_t* b_s = sm + 2 * P * 128;   // P * 128 bytes
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tb, 64); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, 3 * P * 128);
        tma_load_2d(a_s, &a_map, &bar, 0, 0);
        tma_load_2d(a_s + P * 128, &a_map, &bar, 64, 0);
        tma_load_2d(b_s, &b_map, &bar, 0, 0);
        mbar_wait(&bar, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
        auto mk = [](uint32_t lbo) {
            uint64_t d = 0;
            d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
            d |= (uint64_t)(1024 >> 4) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)2 << 61;
            return d;
        };
        const uint64_t hi_a = mk(P * 128), hi_b = mk(P * 128);
        const uint32_t a16 = (smem_u32(a_s) + sa * 128) >> 4, b16 = (smem_u32(b_s) + sb * 128) >> 4;
        for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem, hi_a | (uint64_t)(a16 + kk * 128), hi_b | (uint64_t)(b16 + kk * 128), idesc, kk > 0);
        umma_commit(&done);
    }
    mbar_wait(&done, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_sel = ((uint32_t)(warp & 3) * 32u) << 16;
    for (int cc = 0; cc < 2; ++cc) {
        float v[32];
        tmem_ld32(tmem + lane_sel + cc * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[threadIdx.x * 64 + cc * 32 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    std::vector<__nv_bfloat16> ha(P * 128), hb(P * 64);
    for (int r = 0; r < P; ++r) {
        for (int c = 0; c < 128; ++c) ha[r * 128 + c] = __float2bfloat16((float)((r * 3 + c * 5) % 7 - 3));
        for (int c = 0; c < 64; ++c) hb[r * 64 + c] = __float2bfloat16((float)((r * 5 + c * 3) % 5 - 2));
    }
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap am, bm;
    {
        cuuint32_t es[2] = {1, 1};
        cuuint64_t dims[2] = {128, P}; cuuint64_t str[1] = {256}; cuuint32_t box[2] = {64, P};
        enc(&am, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t dimsb[2] = {64, P}; cuuint64_t strb[1] = {128};
        enc(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimsb, strb, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<float> ho(128 * 64);
    int n_bad_cfg = 0;
    for (int sa = 0; sa < 20; sa += (sa < 9 ? 1 : 5))
        for (int sb = 0; sb < 20; sb += (sb < 9 ? 1 : 6)) {
            k<<<1, 128, 64 * 1024>>>(am, bm, sa, sb, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("sa=%d sb=%d: CUDA error %s\n", sa, sb, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    float ex = 0.f;
                    for (int kk = 0; kk < 64; ++kk)
                        ex += __bfloat162float(ha[(sa + kk) * 128 + m]) * __bfloat162float(hb[(sb + kk) * 64 + n]);
                    if (ho[m * 64 + n] != ex) ++bad;
                }
            if (bad) ++n_bad_cfg;
            printf("sa=%2d sb=%2d: %s (%d mismatches)\n", sa, sb, bad ? "WRONG" : "exact", bad);
        }
    printf("MN-major shifted views: %s\n", n_bad_cfg ? "NOT usable as is" : "all exact");
    return 0;
}
