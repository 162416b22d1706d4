// point_upper.cu -- fused per-frame radar point-cloud encoder of Upper_Net (kernel "K1").
//
// One persistent CTA (128 threads, one point per thread per 128-point chunk) per frame:
//   Transform2H in place (xyz <- R (xyz - t), written back to the caller's tensor: Util/Universal_Util/Utils.py:284-292)
//   -> PointNet 6->8->16->24 (+cat x[0:4])            Net/Upper_Net.py:242-268
//   -> GlobalPointNet 28->32->48->64                   Net/Upper_Net.py:271-297
//   -> attention softmax over the N points + pooling   Net/Upper_Net.py:299-300
// BatchNorm is folded on the host; weights (25 KB) are staged once per CTA in shared memory and read with 128-bit
// broadcast loads; activations never leave registers.  Outputs: g [F,64], global_weights [F,N].
#include "internal.h"
#include "point_layout.h"

namespace mmego {

namespace {

using UL = UpperPointLayout;
constexpr int PT = 128;          // threads per CTA = points per chunk

// The folded MLP weights live in constant memory: every thread of a warp needs the same weight at the same time and the
// layer loops are fully unrolled, so each weight becomes an immediate constant-bank operand of its FFMA -- no load
// instruction at all.  (Reading them from shared memory instead costs one broadcast LDS per four FFMAs and capped the
// kernel at ~22 TFLOP/s.)  Refreshed from the handle's packed blob, stream-ordered, before every launch.
__constant__ float c_w[UL::TOTAL];
constexpr int RED_LD = PT + 1;   // padded row of the transposed reduction tile

template <int CINP, int COUT, bool RELU>
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ b, const float* x,
                                      float* y) {
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
        float a = b[o];
#pragma unroll
        for (int c = 0; c < CINP; ++c) a = fmaf(W[o * CINP + c], x[c], a);
        y[o] = RELU ? fmaxf(a, 0.f) : a;
    }
}

__device__ __forceinline__ float block_reduce_max(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
#pragma unroll
    for (int w = 1; w < PT / 32; ++w) r = fmaxf(r, scratch[w]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
#pragma unroll
    for (int w = 1; w < PT / 32; ++w) r += scratch[w];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(PT) upper_point_kernel(float* __restrict__ x, const float* __restrict__ R,
                                                         const float* __restrict__ t,
                                                         const float* __restrict__ wblob, float* __restrict__ gout,
                                                         float* __restrict__ gw, long long F, int N) {
    MMEGO_DYN_SMEM(float, smem);
    const float* sw = c_w;                     // constant bank
    float* red = smem;                         // [64][RED_LD]
    float* part = red + 64 * RED_LD;           // [2][64]
    float* scratch = part + 128;               // [8]
    float* srt = scratch + 8;                  // [12] R,t of the frame
    const int tid = threadIdx.x;
    (void)wblob;

    for (long long f = blockIdx.x; f < F; f += gridDim.x) {
        if (tid < 9) srt[tid] = R[f * 9 + tid];
        else if (tid < 12) srt[tid] = t[f * 3 + tid - 9];
        __syncthreads();
        float run_m = -INFINITY, run_s = 0.f, run_g = 0.f;   // run_g: pooled channel (tid%64), valid for tid < 64
        float* xf = x + f * (long long)N * 6;
        float* gwf = gw ? gw + f * (long long)N : nullptr;
        for (int p0 = 0; p0 < N; p0 += PT) {
            const int p = p0 + tid;
            const bool live = p < N;
            float psi[64];
            float score = -INFINITY;
            if (live) {
                float in[8];
                float2 v0 = *reinterpret_cast<const float2*>(xf + p * 6);
                float2 v1 = *reinterpret_cast<const float2*>(xf + p * 6 + 2);
                float2 v2 = *reinterpret_cast<const float2*>(xf + p * 6 + 4);
                const float dx = v0.x - srt[9], dy = v0.y - srt[10], dz = v1.x - srt[11];
                in[0] = srt[0] * dx + srt[1] * dy + srt[2] * dz;
                in[1] = srt[3] * dx + srt[4] * dy + srt[5] * dz;
                in[2] = srt[6] * dx + srt[7] * dy + srt[8] * dz;
                in[3] = v1.y; in[4] = v2.x; in[5] = v2.y; in[6] = 0.f; in[7] = 0.f;
                *reinterpret_cast<float2*>(xf + p * 6) = make_float2(in[0], in[1]);
                xf[p * 6 + 2] = in[2];
                float a1[8], a2[16], a3[28];
                dense<8, 8, true>(sw + UL::W1, sw + UL::B1, in, a1);
                dense<8, 16, true>(sw + UL::W2, sw + UL::B2, a1, a2);
                a3[0] = in[0]; a3[1] = in[1]; a3[2] = in[2]; a3[3] = in[3];
                dense<16, 24, true>(sw + UL::W3, sw + UL::B3, a2, a3 + 4);
                float a4[32], a5[48];
                dense<28, 32, true>(sw + UL::W4, sw + UL::B4, a3, a4);
                dense<32, 48, true>(sw + UL::W5, sw + UL::B5, a4, a5);
                dense<48, 64, true>(sw + UL::W6, sw + UL::B6, a5, psi);
                score = sw[UL::BA];
#pragma unroll
                for (int c = 0; c < 64; ++c) score = fmaf(sw[UL::WA + c], psi[c], score);
                if (gwf) gwf[p] = score;       // raw score; normalised after the last chunk
            }
            const float cm = block_reduce_max(score, scratch);
            const float new_m = fmaxf(run_m, cm);
            const float e = live ? expf(score - new_m) : 0.f;
            const float cs = block_reduce_sum(e, scratch);
            const float rescale = (run_m == -INFINITY) ? 0.f : expf(run_m - new_m);
#pragma unroll
            for (int c = 0; c < 64; ++c) red[c * RED_LD + tid] = live ? e * psi[c] : 0.f;
            __syncthreads();
            {
                const int c = tid & 63, half = tid >> 6;
                float s = 0.f;
                const float* rp = red + c * RED_LD + half * 64;
#pragma unroll 16
                for (int i = 0; i < 64; ++i) s += rp[i];
                part[half * 64 + c] = s;
            }
            __syncthreads();
            if (tid < 64) run_g = run_g * rescale + part[tid] + part[64 + tid];
            run_s = run_s * rescale + cs;
            run_m = new_m;
        }
        const float inv = 1.0f / run_s;
        if (tid < 64) gout[f * 64 + tid] = run_g * inv;
        if (gwf)
            for (int p = tid; p < N; p += PT) gwf[p] = expf(gwf[p] - run_m) * inv;
        __syncthreads();   // srt / red reuse
    }
}

}  // namespace

size_t upper_point_smem_bytes() { return (size_t)(64 * RED_LD + 128 + 8 + 12 + 4) * sizeof(float); }

void launch_upper_point(float* x, const float* R, const float* t, const float* wblob, float* g, float* gw,
                        long long F, int N, int sm_count, cudaStream_t st) {
    if (F <= 0) return;
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set)) {
        cudaFuncSetAttribute(upper_point_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)upper_point_smem_bytes());
    }
    cudaMemcpyToSymbolAsync(c_w, wblob, sizeof(float) * UL::TOTAL, 0, cudaMemcpyDeviceToDevice, st);
    long long grid = F < (long long)sm_count * 4 ? F : (long long)sm_count * 4;
    MMEGO_LAUNCH(upper_point_kernel, dim3((unsigned)grid), dim3(PT), upper_point_smem_bytes(), st, x, R, t, wblob, g,
                 gw, F, N);
}

}  // namespace mmego
