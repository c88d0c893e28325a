// Isolated checks: packed fp32x2 arithmetic and a non-swizzled fp32 4-D TMA box load with negative start coordinates.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../licos_b200/csrc/common.cuh"
#include "../licos_b200/csrc/epilogue.cuh"
extern "C" void licos_set_last_cuda_error(int) {}
using namespace licos;

__global__ void k_f2(const float* in, float* out) {
    const uint64_t a = f2_pack(in[0], in[1]), b = f2_pack(in[2], in[3]);
    float x, y;
    f2_unpack(f2_mul(a, b), x, y);
    out[0] = x; out[1] = y;
    f2_unpack(f2_add(a, b), x, y);
    out[2] = x; out[3] = y;
    out[4] = fast_rsqrt(in[0]);
}

__global__ void k_tma(const __grid_constant__ CUtensorMap m, float* out, int c0, int c1, int bw) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bw * 19 * 3 * 4);
        tma_load_4d(sm, &m, &bar, c0, c1, 0, 1);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bw * 19 * 3; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    const int bw = argc > 1 ? atoi(argv[1]) : 36, c0 = argc > 2 ? atoi(argv[2]) : -2;
    float hin[4] = {4.f, 3.f, 0.5f, -2.f}, hout[5];
    float *din, *dout;
    cudaMalloc(&din, 16); cudaMalloc(&dout, 20);
    cudaMemcpy(din, hin, 16, cudaMemcpyHostToDevice);
    k_f2<<<1, 1>>>(din, dout);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hout, dout, 20, cudaMemcpyDeviceToHost);
    printf("f2: %s  mul=(%g,%g) add=(%g,%g) rsqrt=%g\n", cudaGetErrorString(e), hout[0], hout[1], hout[2], hout[3], hout[4]);
    if (e != cudaSuccess) return 1;

    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const int B = 2, C = 3, H = 64, W = 64;
    std::vector<float> hx((size_t)B * C * H * W);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)(i % 9973);
    float* dx; cudaMalloc(&dx, hx.size() * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[4] = {W, H, C, B}; cuuint64_t str[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, 19, 3, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    float* dp; cudaMalloc(&dp, bw * 19 * 3 * 4);
    {
        k_tma<<<1, 128, 16384>>>(m, dp, c0, -2, bw);
        e = cudaDeviceSynchronize();
        std::vector<float> hp(bw * 19 * 3);
        cudaMemcpy(hp.data(), dp, hp.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 19; ++rr) for (int cc = 0; cc < bw; ++cc) {
            const int ih = rr - 2, iw = cc + c0;
            const float ex = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? hx[(((size_t)1 * C + c) * H + ih) * W + iw] : 0.f;
            if (hp[(c * 19 + rr) * bw + cc] != ex) ++bad;
        }
        printf("tma box=%d c0=%d: %s, %d mismatches\n", bw, c0, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
