// mma_frag.cuh -- warp-level m16n8k16 tensor-core fragments for the small per-point / per-frame GEMM chains
// (point_upper.cu, lower_frame.cu, lstm_small.cu).
//
// Why mma.sync and not tcgen05 here: these chains are 3-6 tiny layers (K, N <= 64) applied to 16-point tiles whose
// activations must go layer -> bias/ReLU -> next layer.  With mma.sync the accumulator fragment of layer l IS (after
// packing) the A fragment of layer l+1, so a tile never leaves registers; tcgen05 would bounce every layer through
// TMEM -> registers -> shared memory -> async proxy.  Measured on this pool's B200 (scripts/ubench/mma_rate.cu):
// mma.sync m16n8k16 f16 sustains 557 TFLOP/s (991 FMA/clk/SM) = 7.8x the 71.8 TFLOP/s of FFMA.
//
// fp32-grade results from fp16 inputs (same scheme as lstm_tc.cu): x = hi + lo with hi = fp16(x), lo = fp16(x - hi);
//   a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi                       (dropped term a_lo*w_lo ~ 2^-22 |a w|)
// The big term and the two small terms use SEPARATE fp32 accumulators: the tensor core's accumulate truncates, and a
// truncation of the small-term sum is ~2^-11 smaller in absolute size than one of the big-term sum.
//
// Fragment layouts (PTX ISA, mma.m16n8k16 .f16):  g = lane >> 2, t = lane & 3
//   A (16x16, row):  a0 = A[g][2t..2t+1]  a1 = A[g+8][2t..2t+1]  a2 = A[g][2t+8..2t+9]  a3 = A[g+8][2t+8..2t+9]
//   B (16x8,  col):  b0 = B[2t..2t+1][g]  b1 = B[2t+8..2t+9][g]
//   C (16x8):        c0,c1 = C[g][2t..2t+1]   c2,c3 = C[g+8][2t..2t+1]
// => the C fragments of two neighbouring n-tiles (2s, 2s+1), packed to half2, are the A fragment of k-step s.
#pragma once
#include <cstdint>

#include "cuda_compat.h"

#ifndef MMEGO_EMUL
#include <cuda_fp16.h>
#endif

namespace mmego {
namespace frag {

#ifdef MMEGO_EMUL
// ---- CPU emulation (tests/emul): same fragment semantics, fp16 rounding through _Float16 ----------------------
static inline float h2f_bits(uint16_t b) {
    _Float16 h;
    std::memcpy(&h, &b, 2);
    return (float)h;
}
static inline uint16_t f2h_bits(float x) {
    if (x > 65504.f) x = 65504.f;          // satfinite
    if (x < -65504.f) x = -65504.f;
    _Float16 h = (_Float16)x;
    uint16_t b;
    std::memcpy(&b, &h, 2);
    return b;
}
static inline uint32_t pack_h2(float x, float y) { return (uint32_t)f2h_bits(x) | ((uint32_t)f2h_bits(y) << 16); }
static inline float2 unpack_h2(uint32_t v) { return make_float2(h2f_bits((uint16_t)(v & 0xffffu)), h2f_bits((uint16_t)(v >> 16))); }

// D = A*B + D on one warp: every lane publishes its A and B registers, then evaluates its four outputs.
static inline void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    struct Pub { uint32_t a[4], b[2]; uint32_t pad[2]; } mine, all[32];
    for (int i = 0; i < 4; ++i) mine.a[i] = a[i];
    mine.b[0] = b0; mine.b[1] = b1; mine.pad[0] = mine.pad[1] = 0;
    emul::warp_allgather(&mine, all, sizeof(Pub));
    const int lane = emul::lane_id(), g = lane >> 2, t = lane & 3;
    auto A = [&](int row, int k) {
        const Pub& p = all[(row & 7) * 4 + ((k & 7) >> 1)];
        const float2 v = unpack_h2(p.a[(row >> 3) + ((k >> 3) << 1)]);
        return (k & 1) ? v.y : v.x;
    };
    auto B = [&](int k, int n) {
        const Pub& p = all[n * 4 + ((k & 7) >> 1)];
        const float2 v = unpack_h2(p.b[k >> 3]);
        return (k & 1) ? v.y : v.x;
    };
    for (int i = 0; i < 4; ++i) {
        const int row = g + ((i >> 1) << 3), col = 2 * t + (i & 1);
        double s = 0.0;
        for (int k = 0; k < 16; ++k) s += (double)A(row, k) * (double)B(k, col);
        c[i] = (float)((double)c[i] + s);
    }
}
#else
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));   // first source -> upper half
    return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#endif

// (x, y) -> hi and lo half2 words
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    hi = pack_h2(x, y);
    const float2 h = unpack_h2(hi);
    lo = pack_h2(x - h.x, y - h.y);
}

// One fp16x3 product step: B fragment word order {hi.b0, hi.b1, lo.b0, lo.b1} (as packed by pack.cpp: pack_mma_weight)
__device__ __forceinline__ void mma3(float (&big)[4], float (&small)[4], const uint32_t (&ahi)[4],
                                     const uint32_t (&alo)[4], const uint4& b) {
    mma16816(big, ahi, b.x, b.y);
    mma16816(small, ahi, b.z, b.w);
    mma16816(small, alo, b.x, b.y);
}

// Dense layer on one 16-row tile held in A fragments:  out[j] = act((A W^T) * oscale + bias), n-tile j0 + j = cols
// 8(j0+j) .. +7 of a layer packed with NTW n-tiles in total.
//   wf   : fragment-ordered weights, uint4 index (s*NTW + j0 + j)*32 + lane   (shared memory, or global through L1)
//   bias : float[NTW*8]
template <int KS, int NT, bool RELU, int NTW = NT>
__device__ __forceinline__ void dense_tile(const uint4* __restrict__ wf, const float* __restrict__ bias, float oscale,
                                           const uint32_t (&ahi)[KS][4], const uint32_t (&alo)[KS][4],
                                           float (&out)[NT][4], int lane, int j0 = 0) {
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        float big[4] = {0.f, 0.f, 0.f, 0.f}, small[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < KS; ++s) mma3(big, small, ahi[s], alo[s], wf[(s * NTW + j0 + j) * 32 + lane]);
        const float2 bv = *reinterpret_cast<const float2*>(bias + 8 * (j0 + j) + 2 * t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = fmaf(big[i] + small[i], oscale, (i & 1) ? bv.y : bv.x);
            out[j][i] = RELU ? fmaxf(v, 0.f) : v;
        }
    }
}

// The same layer on TWO independent 16-row tiles of one warp: every B fragment is loaded from shared memory once and
// feeds both tiles, and the two tiles' MMA chains interleave (the layer chains are latency-bound when a warp carries a
// single tile: each layer's MMAs depend on the previous layer's epilogue).
template <int KS, int NT, bool RELU, int NTW = NT>
__device__ __forceinline__ void dense_tile2(const uint4* __restrict__ wf, const float* __restrict__ bias, float oscale,
                                            const uint32_t (&ahi)[2][KS][4], const uint32_t (&alo)[2][KS][4],
                                            float (&out)[2][NT][4], int lane, int j0 = 0) {
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        float big[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, small[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const uint4 b = wf[(s * NTW + j0 + j) * 32 + lane];
            mma3(big[0], small[0], ahi[0][s], alo[0][s], b);
            mma3(big[1], small[1], ahi[1][s], alo[1][s], b);
        }
        const float2 bv = *reinterpret_cast<const float2*>(bias + 8 * (j0 + j) + 2 * t);
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = fmaf(big[u][i] + small[u][i], oscale, (i & 1) ? bv.y : bv.x);
                out[u][j][i] = RELU ? fmaxf(v, 0.f) : v;
            }
    }
}

// C fragments of NT n-tiles -> A fragments of KS k-steps (k-step s = n-tiles 2s, 2s+1; missing tiles are zero)
template <int NT, int KS>
__device__ __forceinline__ void to_afrag(const float (&c)[NT][4], uint32_t (&ahi)[KS][4], uint32_t (&alo)[KS][4]) {
#pragma unroll
    for (int s = 0; s < KS; ++s) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * s + h;
            if (j < NT) {
                split2(c[j][0], c[j][1], ahi[s][2 * h], alo[s][2 * h]);
                split2(c[j][2], c[j][3], ahi[s][2 * h + 1], alo[s][2 * h + 1]);
            } else {
                ahi[s][2 * h] = ahi[s][2 * h + 1] = alo[s][2 * h] = alo[s][2 * h + 1] = 0u;
            }
        }
    }
}

}  // namespace frag
}  // namespace mmego
