// gcn.cu -- the memory-bound glue kernels of the ST-GCN key encoder (Net/GCN.py, Net/Lower_Net.py:149-167).
// Activations are kept channel-last, row = (frame * 15 + joint), so that every convolution of the network is a row-major
// GEMM (gemm_ffma.cu): the 1x1 graph conv after a sparse 15x15 neighbourhood aggregation, and the (9,1) temporal conv as
// nine row-shifted K segments plus the residual 1x1 conv as a tenth.
#include "internal.h"

#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
namespace mmego {

namespace {

// Lower_Net front: Transform2H of the upper-body joints (Net/Lower_Net.py:229) and data_bn (Net/GCN.py:339-341).
//   upper [F,15,3] world -> uh [F,45] head frame (also the 45 extra inputs of fusion.fc0), y0 [F*15,3] normalised.
__global__ void gcn_prep_kernel(const float* __restrict__ upper, const float* __restrict__ R,
                                const float* __restrict__ t, const float* __restrict__ bn, float* __restrict__ uh,
                                float* __restrict__ y0, long long F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (frame, joint)
    if (i >= F * kGcnV) return;
    const long long f = i / kGcnV;
    const int v = (int)(i % kGcnV);
    const float* r = R + f * 9;
    const float* tt = t + f * 3;
    const float* p = upper + i * 3;
    const float dx = p[0] - tt[0], dy = p[1] - tt[1], dz = p[2] - tt[2];
    float h[3];
    h[0] = r[0] * dx + r[1] * dy + r[2] * dz;
    h[1] = r[3] * dx + r[4] * dy + r[5] * dz;
    h[2] = r[6] * dx + r[7] * dy + r[8] * dz;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        uh[i * 3 + c] = h[c];
        y0[i * 3 + c] = h[c] * bn[v * 3 + c] + bn[45 + v * 3 + c];
    }
}

// standalone GCN.Model.extract_feature entry: x [B,3,T,15] -> y0 [(b,t,v), 3] with data_bn applied
__global__ void gcn_prep_raw_kernel(const float* __restrict__ x, const float* __restrict__ bn, float* __restrict__ y0,
                                    int B, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, t, v)
    const long long total = (long long)B * T * kGcnV;
    if (i >= total) return;
    const int v = (int)(i % kGcnV);
    const long long bt = i / kGcnV;
    const int tt = (int)(bt % T);
    const long long b = bt / T;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float val = x[((b * 3 + c) * T + tt) * kGcnV + v];
        y0[i * 3 + c] = val * bn[v * 3 + c] + bn[45 + v * 3 + c];
    }
}

// neighbourhood aggregation: ya[(f,w)][k*C + c] = sum_v y[(f,v)][c] * Ahat[k][v][w]     (einsum of Net/GCN.py:62,
// applied before the 1x1 conv, which commutes with it)
__global__ void gcn_agg_kernel(const float* __restrict__ y, const float* __restrict__ ahat, float* __restrict__ ya,
                               long long F, int C) {
    __shared__ float sa[2 * kGcnV * kGcnV];
    for (int i = threadIdx.x; i < 2 * kGcnV * kGcnV; i += blockDim.x) sa[i] = ahat[i];
    __syncthreads();
    const long long total = F * kGcnV * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long fw = i / C;
        const int w = (int)(fw % kGcnV);
        const long long f = fw / kGcnV;
        const float* yf = y + f * kGcnV * C + c;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int v = 0; v < kGcnV; ++v) {
            const float val = yf[v * C];
            a0 = fmaf(val, sa[v * kGcnV + w], a0);
            a1 = fmaf(val, sa[kGcnV * kGcnV + v * kGcnV + w], a1);
        }
        ya[fw * (2 * C) + c] = a0;
        ya[fw * (2 * C) + C + c] = a1;
    }
}

}  // namespace

void launch_gcn_prep(const float* upper, const float* R, const float* t, const float* bn, float* uh, float* y0,
                     long long F, cudaStream_t st) {
    const long long total = F * kGcnV;
    if (total <= 0) return;
    MMEGO_LAUNCH(gcn_prep_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, upper, R, t, bn, uh, y0, F);
}
void launch_gcn_prep_raw(const float* x, const float* bn, float* y0, int B, int T, cudaStream_t st) {
    const long long total = (long long)B * T * kGcnV;
    if (total <= 0) return;
    MMEGO_LAUNCH(gcn_prep_raw_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, x, bn, y0, B, T);
}
void launch_gcn_agg(const float* y, const float* ahat, float* ya, long long F, int C, int sm_count, cudaStream_t st) {
    const long long total = F * kGcnV * C;
    if (total <= 0) return;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count * 16;
    if (blocks > cap) blocks = cap;
    MMEGO_LAUNCH(gcn_agg_kernel, dim3((unsigned)blocks), dim3(256), 0, st, y, ahat, ya, F, C);
}

}  // namespace mmego
#endif  // MMEGO_FFMA_GEN
