// MS-SSIM on the device (/root/reference/eval_utils.py:159-169 `compute_msssim` =
// pytorch_msssim.ms_ssim(a, b, data_range=1.0)): per pyramid level a separable 11-tap Gaussian (sigma 1.5, "valid"
// borders) of x, y, x^2, y^2, xy; SSIM and contrast-structure maps; their means per (image, channel); 2x2 average pooling
// to the next level.  HBM-bound stencil work: the horizontal pass writes five filtered planes, the vertical pass
// consumes them, forms the maps and reduces them in the same kernel (double accumulation, one atomic per block).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/licos_b200.h"
#include "common.cuh"

namespace licos {

constexpr int kWin = 11;
struct GaussWin { float w[kWin]; };

// planes: [5][BC][H][Wo], Wo = W - 10
__global__ void msssim_h_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t rows, int W, int Wo,
                                GaussWin g, float* __restrict__ planes) {
    const int64_t total = rows * Wo;
    const int64_t plane = total;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / Wo;
        const int c = (int)(e - r * Wo);
        const float* xr = x + r * W + c;
        const float* yr = y + r * W + c;
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float a = __ldg(xr + k), b = __ldg(yr + k), w = g.w[k];
            sx += w * a; sy += w * b; sxx += w * a * a; syy += w * b * b; sxy += w * a * b;
        }
        planes[e] = sx; planes[plane + e] = sy; planes[2 * plane + e] = sxx; planes[3 * plane + e] = syy; planes[4 * plane + e] = sxy;
    }
}

// sums[bc][0] += sum(ssim_map), sums[bc][1] += sum(cs_map) over the (H-10) x Wo valid positions of image-channel bc
__global__ void msssim_v_kernel(const float* __restrict__ planes, int BC, int H, int Wo, GaussWin g, float C1, float C2,
                                double* __restrict__ sums) {
    const int Ho = H - (kWin - 1);
    const int bc = blockIdx.y;
    const int64_t plane = (int64_t)BC * H * Wo;
    const float* base = planes + (int64_t)bc * H * Wo;
    double acc_s = 0.0, acc_c = 0.0;
    const int64_t total = (int64_t)Ho * Wo;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / Wo), c = (int)(e - (int64_t)r * Wo);
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const int64_t o = (int64_t)(r + k) * Wo + c;
            const float w = g.w[k];
            m1 += w * __ldg(base + o); m2 += w * __ldg(base + plane + o); s11 += w * __ldg(base + 2 * plane + o);
            s22 += w * __ldg(base + 3 * plane + o); s12 += w * __ldg(base + 4 * plane + o);
        }
        const float mu11 = m1 * m1, mu22 = m2 * m2, mu12 = m1 * m2;
        const float v1 = s11 - mu11, v2 = s22 - mu22, v12 = s12 - mu12;
        const float cs = (2.f * v12 + C2) / (v1 + v2 + C2);
        const float ss = ((2.f * mu12 + C1) / (mu11 + mu22 + C1)) * cs;
        acc_s += (double)ss;
        acc_c += (double)cs;
    }
    __shared__ double red[2][32];
    for (int o = 16; o > 0; o >>= 1) {
        acc_s += __shfl_down_sync(0xffffffffu, acc_s, o);
        acc_c += __shfl_down_sync(0xffffffffu, acc_c, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { red[0][w] = acc_s; red[1][w] = acc_c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; b += red[1][i]; }
        atomicAdd(sums + 2 * bc, a);
        atomicAdd(sums + 2 * bc + 1, b);
    }
}

// F.avg_pool2d(x, kernel_size=2, padding=(H % 2, W % 2)) with count_include_pad: out = (sum of the 2x2 window) / 4
__global__ void avgpool2_kernel(const float* __restrict__ x, int64_t BC, int H, int W, int Ho, int Wo, int ph, int pw,
                                float* __restrict__ out) {
    const int64_t total = BC * Ho * Wo;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % Wo);
        const int r = (int)((e / Wo) % Ho);
        const int64_t bc = e / ((int64_t)Wo * Ho);
        const float* img = x + bc * H * W;
        float s = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int ih = 2 * r + dy - ph, iw = 2 * c + dx - pw;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) s += __ldg(img + (int64_t)ih * W + iw);
            }
        out[e] = 0.25f * s;
    }
}

// Raw Sentinel-2 digital numbers -> model input (/root/reference/licos/raw_image_folder.py:192-196, raw_utils.py:128):
// band = dn / 4095 (float64); unless use_full_range: band = rint(band * 255) / 255 (skimage.img_as_ubyte); float32.
__global__ void raw_dn_kernel(const uint16_t* __restrict__ dn, int64_t n, double inv_max, int requant8, float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        double v = (double)dn[e] * inv_max;
        if (requant8) v = rint(v * 255.0) / 255.0;
        out[e] = (float)v;
    }
}

static int grid_cap(int64_t n) {
    int64_t g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace licos

using namespace licos;

extern "C" {

int64_t licos_msssim_workspace_floats(int64_t bc, int h, int w) {
    if (bc < 0 || h < kWin || w < kWin) return LICOS_ERR_INVALID;
    return 5 * bc * (int64_t)h * (w - (kWin - 1));
}

int licos_msssim_level(const float* x, const float* y, int64_t bc, int h, int w, const float* win11, float c1, float c2,
                       float* workspace, double* sums, void* stream) {
    if (!x || !y || !win11 || !workspace || !sums || bc < 0 || h < kWin || w < kWin || bc > 0x7fffffff) return LICOS_ERR_INVALID;
    if (bc == 0) return LICOS_OK;
    GaussWin g;
    for (int i = 0; i < kWin; ++i) g.w[i] = win11[i];  // host pointer: 11 taps
    cudaStream_t s = (cudaStream_t)stream;
    const int wo = w - (kWin - 1), ho = h - (kWin - 1);
    msssim_h_kernel<<<grid_cap(bc * h * wo), 256, 0, s>>>(x, y, bc * h, w, wo, g, workspace);
    LICOS_CUDA_OK(cudaGetLastError());
    int gx = (int)(((int64_t)ho * wo + 255) / 256);
    const int cap = (int)((148 * 16 + bc - 1) / bc);
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    if (bc > 65535) return LICOS_ERR_UNSUPPORTED;
    msssim_v_kernel<<<dim3(gx, (unsigned)bc), 256, 0, s>>>(workspace, (int)bc, h, wo, g, c1, c2, sums);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_raw_dn_to_unit(const uint16_t* dn, int64_t n, int dn_max, int requant8, float* out, void* stream) {
    if (!dn || !out || n < 0 || dn_max < 1) return LICOS_ERR_INVALID;
    if (n == 0) return LICOS_OK;
    raw_dn_kernel<<<grid_cap(n), 256, 0, (cudaStream_t)stream>>>(dn, n, 1.0 / (double)dn_max, requant8, out);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

int licos_avgpool2(const float* x, int64_t bc, int h, int w, float* out, void* stream) {
    if (!x || !out || bc < 0 || h < 1 || w < 1) return LICOS_ERR_INVALID;
    if (bc == 0) return LICOS_OK;
    const int ph = h % 2, pw = w % 2;
    const int ho = (h + 2 * ph - 2) / 2 + 1, wo = (w + 2 * pw - 2) / 2 + 1;
    avgpool2_kernel<<<grid_cap(bc * ho * wo), 256, 0, (cudaStream_t)stream>>>(x, bc, h, w, ho, wo, ph, pw, out);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

}  // extern "C"
