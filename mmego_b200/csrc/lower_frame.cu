// lower_frame.cu -- fused per-frame front end of Lower_Net (kernels "K1'" + the cross-attention part of "K3").
//
// One persistent CTA (128 threads) per frame:
//   second in-place Transform2H of the cloud (Net/Lower_Net.py:191-192; the cloud was already transformed once by
//     UpperNet.forward, and the trained weights expect exactly that)
//   -> top-64 of N points by transformed x (Net/Lower_Net.py:216-227) as a rank select: a point's slot is the number
//      of points that beat it; equal keys: the lower slot index wins (documented tie rule)
//   -> BasePointNet 6->16->32->61, cat xyz -> P [64,64]                         Net/Lower_Net.py:40-72
//   -> to_q on P, to_k / to_v on the frame's 15 ST-GCN joint features, softmax(q k^T / 8) over joints   :104-108
//   -> a = sum over the 64 points of [P | attn @ v]   (the reference's second "attention" softmaxes a size-1
//      dimension, i.e. all weights are 1: Net/Lower_Net.py:111-113), kbar = mean over joints of K       :114-115
// Output: ak [F,192] = [a (128) | kbar (64)], the input of rnn_pk.
// sum_s (alpha[s,:] @ v) is evaluated as (sum_s alpha[s,:]) @ v, so attn @ v is never materialised.
#include "internal.h"
#include "point_layout.h"

namespace mmego {

namespace {

using LL = LowerFrameLayout;
constexpr int NT = 128;
// BasePointNet's folded weights (W1..B3, 2.8k floats) in constant memory: immediate FFMA operands, as in point_upper.cu.
// The three 64x64 projections stay in shared memory (they are indexed per thread).
constexpr int kMlpFloats = LL::WQ;
__constant__ float c_mlp[kMlpFloats];
constexpr int NMAX = 512;      // max points per frame supported by the rank select
constexpr int LDP = 65;

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

struct Smem {
    float w[LL::TOTAL];
    float pts[NMAX * 6];
    float key[NMAX];
    int sel[kLowerPts];
    float P[kLowerPts * LDP];
    float Kf[kGcnV * 64];
    float kp[kGcnV * 64];      // to_k(K)
    float vp[kGcnV * 64];      // to_v(K)
    float spart[2 * kLowerPts * 16];
    float asum[2 * 16];
    float rt[12];
};

__global__ void __launch_bounds__(NT) lower_frame_kernel(float* __restrict__ x, const float* __restrict__ R,
                                                         const float* __restrict__ t,
                                                         const float* __restrict__ kfeat,
                                                         const float* __restrict__ wblob, float* __restrict__ ak,
                                                         long long F, int N) {
    MMEGO_DYN_SMEM(Smem, sp);
    Smem& s = *sp;
    const int tid = threadIdx.x;
    for (int i = tid * 4; i < LL::TOTAL; i += NT * 4)
        *reinterpret_cast<float4*>(s.w + i) = *reinterpret_cast<const float4*>(wblob + i);

    for (long long f = blockIdx.x; f < F; f += gridDim.x) {
        __syncthreads();   // previous frame fully consumed (and weights staged on the first pass)
        if (tid < 9) s.rt[tid] = R[f * 9 + tid];
        else if (tid < 12) s.rt[tid] = t[f * 3 + tid - 9];
        for (int i = tid; i < kGcnV * 64; i += NT) s.Kf[i] = kfeat[f * (kGcnV * 64) + i];
        __syncthreads();
        // ---- A: second transform (in place) + keys -----------------------------------------------------
        float* xf = x + f * (long long)N * 6;
        for (int p = tid; p < N; p += NT) {
            float2 v0 = *reinterpret_cast<const float2*>(xf + p * 6);
            float2 v1 = *reinterpret_cast<const float2*>(xf + p * 6 + 2);
            float2 v2 = *reinterpret_cast<const float2*>(xf + p * 6 + 4);
            const float dx = v0.x - s.rt[9], dy = v0.y - s.rt[10], dz = v1.x - s.rt[11];
            const float nx = s.rt[0] * dx + s.rt[1] * dy + s.rt[2] * dz;
            const float ny = s.rt[3] * dx + s.rt[4] * dy + s.rt[5] * dz;
            const float nz = s.rt[6] * dx + s.rt[7] * dy + s.rt[8] * dz;
            *reinterpret_cast<float2*>(xf + p * 6) = make_float2(nx, ny);
            xf[p * 6 + 2] = nz;
            float* pp = s.pts + p * 6;
            pp[0] = nx; pp[1] = ny; pp[2] = nz; pp[3] = v1.y; pp[4] = v2.x; pp[5] = v2.y;
            s.key[p] = nx;
        }
        __syncthreads();
        for (int p = tid; p < N; p += NT) {
            const float kx = s.key[p];
            int rank = 0;
            for (int q = 0; q < N; ++q) {
                const float kq = s.key[q];
                rank += (kq > kx || (kq == kx && q < p)) ? 1 : 0;
            }
            if (rank < kLowerPts) s.sel[rank] = p;
        }
        __syncthreads();
        // ---- B: per-point MLP (threads 0..63) || to_k / to_v of the joint features (threads 64..127) ----
        if (tid < kLowerPts) {
            const float* pp = s.pts + s.sel[tid] * 6;
            float in[8] = {pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], 0.f, 0.f};
            float a1[16], a2[32], a3[64];
            dense<8, 16, true>(c_mlp + LL::W1, c_mlp + LL::B1, in, a1);
            dense<16, 32, true>(c_mlp + LL::W2, c_mlp + LL::B2, a1, a2);
            dense<32, 64, true>(c_mlp + LL::W3, c_mlp + LL::B3, a2, a3);   // rows 61..63 are zero padding
            float* pr = s.P + tid * LDP;
            pr[0] = in[0]; pr[1] = in[1]; pr[2] = in[2];
#pragma unroll
            for (int c = 0; c < 61; ++c) pr[3 + c] = a3[c];
        } else {
            const int o = tid - kLowerPts;   // output channel
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                const float* WT = s.w + (which ? LL::WV : LL::WK);   // transposed: [c][o]
                const float bias = s.w[(which ? LL::BV : LL::BK) + o];
                float* dst = which ? s.vp : s.kp;
                float wc[64];
#pragma unroll
                for (int c = 0; c < 64; ++c) wc[c] = WT[c * 64 + o];
#pragma unroll 1
                for (int j = 0; j < kGcnV; ++j) {
                    float a = bias;
#pragma unroll
                    for (int c = 0; c < 64; c += 4) {
                        const float4 kv = *reinterpret_cast<const float4*>(s.Kf + j * 64 + c);
                        a = fmaf(wc[c], kv.x, a);
                        a = fmaf(wc[c + 1], kv.y, a);
                        a = fmaf(wc[c + 2], kv.z, a);
                        a = fmaf(wc[c + 3], kv.w, a);
                    }
                    dst[j * 64 + o] = a;
                }
            }
        }
        __syncthreads();
        // ---- C: to_q on P (each thread: one point, 32 of the 64 outputs) + partial attention scores ----
        {
            const int pt = tid & 63, half = tid >> 6;
            float prow[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) prow[c] = s.P[pt * LDP + c];
            float sc[kGcnV];
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) sc[j] = 0.f;
#pragma unroll 4
            for (int oo = 0; oo < 32; ++oo) {
                const int o = half * 32 + oo;
                float a = s.w[LL::BQ + o];
                const float* wq = s.w + LL::WQ + o * 64;
#pragma unroll
                for (int c = 0; c < 64; c += 4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wq + c);
                    a = fmaf(wv.x, prow[c], a);
                    a = fmaf(wv.y, prow[c + 1], a);
                    a = fmaf(wv.z, prow[c + 2], a);
                    a = fmaf(wv.w, prow[c + 3], a);
                }
#pragma unroll
                for (int j = 0; j < kGcnV; ++j) sc[j] = fmaf(a, s.kp[j * 64 + o], sc[j]);
            }
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) s.spart[(half * kLowerPts + pt) * 16 + j] = sc[j];
        }
        __syncthreads();
        // ---- D: softmax over joints + column sums of alpha (threads 0..63) || column sums of P (64..127) --
        if (tid < kLowerPts) {
            float sc[kGcnV];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) {
                sc[j] = (s.spart[tid * 16 + j] + s.spart[(kLowerPts + tid) * 16 + j]) * 0.125f;
                m = fmaxf(m, sc[j]);
            }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) { sc[j] = expf(sc[j] - m); sum += sc[j]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) {
                float a = sc[j] * inv;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if ((tid & 31) == 0) s.asum[(tid >> 5) * 16 + j] = a;
            }
        } else {
            const int c = tid - kLowerPts;
            float a = 0.f;
#pragma unroll 16
            for (int q = 0; q < kLowerPts; ++q) a += s.P[q * LDP + c];
            ak[f * 192 + c] = a;
        }
        __syncthreads();
        // ---- E: a_T = (sum_s alpha) @ v ; kbar -----------------------------------------------------------
        if (tid < 64) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) a = fmaf(s.asum[j] + s.asum[16 + j], s.vp[j * 64 + tid], a);
            ak[f * 192 + 64 + tid] = a;
        } else {
            const int c = tid - 64;
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) a += s.Kf[j * 64 + c];
            ak[f * 192 + 128 + c] = a * (1.0f / 15.0f);
        }
    }
}

}  // namespace

size_t lower_frame_smem_bytes() { return sizeof(Smem); }
int lower_frame_max_points() { return NMAX; }

void launch_lower_frame(float* x, const float* R, const float* t, const float* kfeat, const float* wblob, float* ak,
                        long long F, int N, int sm_count, cudaStream_t st) {
    if (F <= 0) return;
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set)) {
        cudaFuncSetAttribute(lower_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    }
    cudaMemcpyToSymbolAsync(c_mlp, wblob, sizeof(float) * kMlpFloats, 0, cudaMemcpyDeviceToDevice, st);
    long long grid = F < (long long)sm_count * 2 ? F : (long long)sm_count * 2;
    MMEGO_LAUNCH(lower_frame_kernel, dim3((unsigned)grid), dim3(NT), sizeof(Smem), st, x, R, t, kfeat, wblob, ak, F, N);
}

}  // namespace mmego
