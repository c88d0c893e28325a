// Register-level pieces of the fused conv epilogue shared by the conv kernels: packed fp32x2 arithmetic
// (sm_100 FADD2 / FMUL2), TMEM -> bf16 staging conversion, and the two stages of the fused GDN / IGDN
// (SURVEY.md section 8a row A5):
//   stage 1   x = acc (+ bias);   x kept in registers as packed bf16;   x^2 (bf16) -> swizzled smem tile that is the
//             A operand of the gamma GEMM, which overwrites the accumulator in place with gamma . x^2
//   stage 2   d = norm + beta;    out = x * rsqrt(d)  (GDN)   or   x * d * rsqrt(d) = x * sqrt(d)  (IGDN)
#pragma once

#include "common.cuh"

namespace licos {

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// bf16x2 from an fp32 pair (lo -> low half)
__device__ __forceinline__ uint32_t f2_to_bf16x2(uint64_t v) {
    float lo, hi;
    f2_unpack(v, lo, hi);
    return pack_bf16x2(lo, hi);
}
// fp32 pair from a bf16x2 (exact)
__device__ __forceinline__ uint64_t bf16x2_to_f2(uint32_t v) {
    return f2_pack(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

__device__ __forceinline__ float fast_rsqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 16 packed bf16 pairs (32 consecutive channels of one row) -> the four 16-byte chunks of the swizzled tile
__device__ __forceinline__ void store_row32(uint8_t* tile_base, int row, int col32, const uint32_t (&pk)[16]) {
    uint8_t* atom = tile_base + (size_t)((col32 * 32) / 64) * (128 * 128);
    const uint32_t chunk0 = ((col32 * 32) % 64) / 8;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(atom + sw128_offset(row, chunk0 + q)) =
            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// stage 1 for 32 channels: v (fp32 accumulator columns, bias optional) -> xs (packed x), sq (packed x^2)
template <bool ADD_BIAS>
__device__ __forceinline__ void gdn_stage1_32(const float (&v)[32], const float* bias32, uint32_t* xs16, uint32_t (&sq)[16]) {
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(bias32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint64_t x0 = f2_pack(v[4 * q], v[4 * q + 1]), x1 = f2_pack(v[4 * q + 2], v[4 * q + 3]);
        if (ADD_BIAS) {
            const ulonglong2 b = b2[q];
            x0 = f2_add(x0, b.x);
            x1 = f2_add(x1, b.y);
        }
        xs16[2 * q] = f2_to_bf16x2(x0);
        xs16[2 * q + 1] = f2_to_bf16x2(x1);
        sq[2 * q] = f2_to_bf16x2(f2_mul(x0, x0));
        sq[2 * q + 1] = f2_to_bf16x2(f2_mul(x1, x1));
    }
}

// stage 2 for 32 channels: v = norm columns, beta32 in smem, xs16 = packed x -> out (packed bf16)
template <bool INVERSE>
__device__ __forceinline__ void gdn_stage2_32(const float (&v)[32], const float* beta32, const uint32_t* xs16,
                                              uint32_t (&out)[16]) {
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(beta32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const ulonglong2 b = b2[q];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t d = f2_add(f2_pack(v[4 * q + 2 * h], v[4 * q + 2 * h + 1]), h ? b.y : b.x);
            float d0, d1;
            f2_unpack(d, d0, d1);
            uint64_t r = f2_pack(fast_rsqrt(d0), fast_rsqrt(d1));
            if (INVERSE) r = f2_mul(r, d);  // d * rsqrt(d) = sqrt(d)
            out[2 * q + h] = f2_to_bf16x2(f2_mul(bf16x2_to_f2(xs16[2 * q + h]), r));
        }
    }
}

// The gamma GEMM of the fused GDN: D[tmem] = A (x^2 staging tile, K-major, 128-byte swizzle, atoms of 64 channels)
// x B (resident gamma, same layout), KS = C / 16 MMAs issued back to back.  Call from ONE elected lane of a converged
// warp: inside an elect_one() block the compiler keeps the operands in uniform registers (no per-MMA R2UR / ELECT loop).
template <int KS>
__device__ __forceinline__ void issue_gamma_gemm(uint32_t d_tmem, uint32_t stg16, uint32_t gamma16, uint32_t n_rows,
                                                 uint32_t idesc) {
    const uint64_t desc_hi = umma_desc_sw128(0);
#pragma unroll
    for (uint32_t ks = 0; ks < (uint32_t)KS; ++ks) {
        const uint32_t atom = ks >> 2, off = (ks & 3) * 2;
        umma_bf16(d_tmem, desc_hi | (uint64_t)(stg16 + atom * 1024 + off), desc_hi | (uint64_t)(gamma16 + atom * (n_rows * 8) + off),
                  idesc, (uint32_t)(ks > 0));
    }
}
__device__ __forceinline__ void issue_gamma_gemm_n(uint32_t d_tmem, uint32_t stg16, uint32_t gamma16, uint32_t N, uint32_t idesc) {
    switch (N) {
        case 64: issue_gamma_gemm<4>(d_tmem, stg16, gamma16, N, idesc); break;
        case 128: issue_gamma_gemm<8>(d_tmem, stg16, gamma16, N, idesc); break;
        case 192: issue_gamma_gemm<12>(d_tmem, stg16, gamma16, N, idesc); break;
        default: issue_gamma_gemm<16>(d_tmem, stg16, gamma16, N, idesc); break;  // 256
    }
}

// The same GEMM on a CTA pair (cta_group::2, M = 256): each CTA's staging tile is its 128 rows of A, each CTA holds
// N / 2 rows of gamma (atoms of [N / 2][64]).  Issued by one elected lane of the LEADER CTA.
template <int KS>
__device__ __forceinline__ void issue_gamma_gemm_pair(uint32_t d_tmem, uint32_t stg16, uint32_t gamma16, uint32_t half_rows,
                                                      uint32_t idesc) {
    const uint64_t desc_hi = umma_desc_sw128(0);
#pragma unroll
    for (uint32_t ks = 0; ks < (uint32_t)KS; ++ks) {
        const uint32_t atom = ks >> 2, off = (ks & 3) * 2;
        umma_bf16_pair(d_tmem, desc_hi | (uint64_t)(stg16 + atom * 1024 + off),
                       desc_hi | (uint64_t)(gamma16 + atom * (half_rows * 8) + off), idesc, (uint32_t)(ks > 0));
    }
}
__device__ __forceinline__ void issue_gamma_gemm_pair_n(uint32_t d_tmem, uint32_t stg16, uint32_t gamma16, uint32_t N,
                                                        uint32_t idesc) {
    switch (N) {
        case 64: issue_gamma_gemm_pair<4>(d_tmem, stg16, gamma16, N / 2, idesc); break;
        case 128: issue_gamma_gemm_pair<8>(d_tmem, stg16, gamma16, N / 2, idesc); break;
        case 192: issue_gamma_gemm_pair<12>(d_tmem, stg16, gamma16, N / 2, idesc); break;
        default: issue_gamma_gemm_pair<16>(d_tmem, stg16, gamma16, N / 2, idesc); break;  // 256
    }
}

}  // namespace licos
