// Federated weight merge over NCCL / NVLink (BASELINE.json configs[4]; replaces the file-based merge of
// /root/reference/licos/federation_utils.py:27-85, whose rule is theta <- w_local theta_local + w_central theta_central
// with weights inversely proportional to the losses, :47-53).
//
// All floating-point state of a rank lives in ONE flat fp32 buffer with one spare element at the end.  A merge is
//   1. prep kernel      u = 1 / loss (device scalar, read from device memory: no host sync), flat[n] = 1
//   2. ONE ncclAllReduce over n + 1 elements with a PreMulSum operator whose scalar is u, dereferenced on the device
//      while the collective runs:   flat <- sum_r u_r theta_r,   flat[n] <- sum_r u_r
//   3. normalise kernel flat[0..n) *= 1 / flat[n]   ->  sum_r (u_r / sum_s u_s) theta_r, the N-way form of the reference rule
// i.e. no separate scaling pass before the collective, no all-gather of the weights, nothing read back by the host.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries -- torch's -- else the system one), so
// the library has no link-time dependency on it and the single-GPU paths never touch it.
#include "common.cuh"

#include <dlfcn.h>
#include <math.h>
#include <string.h>
#include <mutex>

namespace licos {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.11: PreMulSum)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclRedOp_t;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclScalarDevice = 0 };

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*RedOpCreatePreMulSum)(ncclRedOp_t*, void*, int, int, ncclComm_t);
    ncclResult_t (*RedOpDestroy)(ncclRedOp_t, ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GetVersion)(int*);
    bool ok;
};

static const NcclApi& nccl() {
    static NcclApi api{};
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy torch.distributed loaded
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.RedOpCreatePreMulSum = (decltype(api.RedOpCreatePreMulSum))dlsym(h, "ncclRedOpCreatePreMulSum");
        api.RedOpDestroy = (decltype(api.RedOpDestroy))dlsym(h, "ncclRedOpDestroy");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.RedOpCreatePreMulSum && api.RedOpDestroy &&
                 api.AllReduce;
    });
    return api;
}

// u = 1 / loss; a rank whose loss is not a positive finite number keeps a vanishing (not zero: all ranks may be in that
// state) weight instead of poisoning every replica with NaN
__global__ void fed_prep_kernel(const float* __restrict__ loss, const float* __restrict__ weight, float* __restrict__ flat,
                                int64_t n, float* __restrict__ u) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float w;
        if (weight) {
            w = *weight;
        } else {
            const float l = *loss;
            w = (isfinite(l) && l > 0.f) ? 1.f / fmaxf(l, 1e-12f) : 1e-20f;
        }
        if (!isfinite(w) || w <= 0.f) w = 1e-20f;
        *u = w;
        flat[n] = 1.f;
    }
}

__global__ void fed_normalise_kernel(float* __restrict__ flat, int64_t n) {
    const float inv = 1.f / flat[n];
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    float4* f4 = reinterpret_cast<float4*>(flat);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = f4[i];
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        f4[i] = v;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) flat[i] *= inv;
}

}  // namespace licos

using namespace licos;

extern "C" {

int licos_nccl_version(void) {
    int v = 0;
    if (!nccl().ok || !nccl().GetVersion || nccl().GetVersion(&v) != kNcclSuccess) return LICOS_ERR_UNSUPPORTED;
    return v;
}

int licos_nccl_unique_id(void* id128_host) {
    if (!id128_host) return LICOS_ERR_INVALID;
    if (!nccl().ok) return LICOS_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != kNcclSuccess) return LICOS_ERR_CUDA;
    memcpy(id128_host, id.internal, 128);
    return LICOS_OK;
}

int licos_nccl_comm_create(const void* id128_host, int world, int rank, void** comm_out) {
    if (!id128_host || !comm_out || world < 1 || rank < 0 || rank >= world) return LICOS_ERR_INVALID;
    if (!nccl().ok) return LICOS_ERR_UNSUPPORTED;
    ncclUniqueId id;
    memcpy(id.internal, id128_host, 128);
    ncclComm_t comm = nullptr;
    if (nccl().CommInitRank(&comm, world, id, rank) != kNcclSuccess) return LICOS_ERR_CUDA;
    *comm_out = comm;
    return LICOS_OK;
}

int licos_nccl_comm_destroy(void* comm) {
    if (!comm) return LICOS_OK;
    if (!nccl().ok) return LICOS_ERR_UNSUPPORTED;
    return nccl().CommDestroy((ncclComm_t)comm) == kNcclSuccess ? LICOS_OK : LICOS_ERR_CUDA;
}

int licos_nccl_weighted_allreduce(void* comm, float* flat, int64_t n, const float* loss_dev, const float* weight_dev,
                                  float* scalar_dev, void* stream) {
    if (!comm || !flat || n < 1 || (!loss_dev && !weight_dev) || !scalar_dev) return LICOS_ERR_INVALID;
    if (((uintptr_t)flat & 15) != 0) return LICOS_ERR_INVALID;
    if (!nccl().ok) return LICOS_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    fed_prep_kernel<<<1, 32, 0, s>>>(loss_dev, weight_dev, flat, n, scalar_dev);
    LICOS_CUDA_OK(cudaGetLastError());
    ncclRedOp_t op;
    if (nccl().RedOpCreatePreMulSum(&op, scalar_dev, kNcclFloat32, kNcclScalarDevice, (ncclComm_t)comm) != kNcclSuccess)
        return LICOS_ERR_CUDA;
    const ncclResult_t r = nccl().AllReduce(flat, flat, (size_t)(n + 1), kNcclFloat32, op, (ncclComm_t)comm, s);
    nccl().RedOpDestroy(op, (ncclComm_t)comm);
    if (r != kNcclSuccess) return LICOS_ERR_CUDA;
    int64_t g = (n / 4 + 255) / 256;
    if (g > 148 * 4) g = 148 * 4;
    if (g < 1) g = 1;
    fed_normalise_kernel<<<(int)g, 256, 0, s>>>(flat, n);
    LICOS_CUDA_OK(cudaGetLastError());
    return LICOS_OK;
}

}  // extern "C"
