// Experiment: can a SWIZZLE_128B K-major UMMA operand start at an arbitrary 128-byte row of a TMA-written tile,
// and can the stride between 8-row groups (SBO) be something other than a multiple of 1024 bytes?
// A = rows [P][64] bf16 loaded by TMA (SWIZZLE_128B) into 1024-aligned smem, B = identity (N = 64, K = 64), so
// D[m][n] must equal A[row(m)][n] with row(m) = s + (m / 8) * (SBO / 128) + m % 8.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

constexpr int P = 256;  // rows in smem

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                                            int s, int sbo_bytes, int base_off, float* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar, done;
    __shared__ uint32_t tb;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* a_s = sm;               // P * 128 bytes
    uint8_t* b_s = sm + P * 128;     // 64 * 128 bytes
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tb, 64); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, P * 128 + 64 * 128);
        tma_load_2d(a_s, &a_map, &bar, 0, 0);
        tma_load_2d(b_s, &b_map, &bar, 0, 0);
        mbar_wait(&bar, 0);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        uint64_t hi_a = 0;
        hi_a |= (uint64_t)1 << 16;
        hi_a |= (uint64_t)(sbo_bytes >> 4) << 32;
        hi_a |= (uint64_t)1 << 46;
        hi_a |= (uint64_t)(base_off & 7) << 49;
        hi_a |= (uint64_t)2 << 61;
        const uint64_t hi_b = umma_desc_sw128(0);
        const uint32_t a16 = (smem_u32(a_s) + s * 128) >> 4, b16 = smem_u32(b_s) >> 4;
        for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem, hi_a | (uint64_t)(a16 + 2 * kk), hi_b | (uint64_t)(b16 + 2 * kk), idesc, kk > 0);
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
    std::vector<__nv_bfloat16> ha(P * 64), hb(64 * 64);
    for (int r = 0; r < P; ++r)
        for (int c = 0; c < 64; ++c) ha[r * 64 + c] = __float2bfloat16((float)(r * 64 + c) / 64.f);  // exact in bf16? use small ints
    for (int r = 0; r < P; ++r)
        for (int c = 0; c < 64; ++c) ha[r * 64 + c] = __float2bfloat16((float)((r * 3 + c * 5) % 251));
    for (int n = 0; n < 64; ++n)
        for (int c = 0; c < 64; ++c) hb[n * 64 + c] = __float2bfloat16(n == c ? 1.f : 0.f);
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap am, bm;
    {
        cuuint64_t dims[2] = {64, P}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, P}; cuuint32_t es[2] = {1, 1};
        enc(&am, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t dimsb[2] = {64, 64}; cuuint32_t boxb[2] = {64, 64};
        enc(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<float> ho(128 * 64);
    const int sbos[] = {1024, 1280, 1152, 2048, 2304};
    for (int sbo : sbos)
        for (int s = 0; s < 10; ++s)
            for (int mode = 0; mode < 2; ++mode) {
                const int bo = mode ? (s & 7) : 0;
                k<<<1, 128, 64 * 1024>>>(am, bm, s, sbo, bo, dout);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sbo=%d s=%d bo=%d: CUDA error %s\n", sbo, s, bo, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0, first_bad = -1;
                for (int m = 0; m < 128; ++m) {
                    const int row = s + (m / 8) * (sbo / 128) + m % 8;
                    if (row >= P) continue;
                    for (int c = 0; c < 64; ++c) {
                        const float ex = __bfloat162float(ha[row * 64 + c]);
                        if (ho[m * 64 + c] != ex) { ++bad; if (first_bad < 0) first_bad = m * 64 + c; }
                    }
                }
                printf("sbo=%4d s=%d base_off=%d: %s (%d mismatches, first at m=%d c=%d)\n", sbo, s, bo, bad ? "WRONG" : "exact", bad,
                       first_bad / 64, first_bad % 64);
            }
    return 0;
}
