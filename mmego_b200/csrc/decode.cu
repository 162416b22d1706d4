// decode.cu -- the HBM-bound "K4" kernels: 6D->rotation, forward kinematics, rigid transforms, IMU attention pooling,
// result assembly and error metrics.  All are streaming kernels: tiles of frames are staged through shared memory
// with coalesced 128-bit (or widest legal) accesses; per-frame math runs one frame per thread out of shared memory.
#include "internal.h"
#include "decode_math.cuh"

namespace mmego {

namespace {

using namespace dec;   // skeleton tables, ortho6d, frame decode / assembly / metrics (decode_math.cuh)

// cooperative tile copy global -> shared (n floats, base 16B-aligned when n_total is), vectorised where possible
__device__ __forceinline__ void tile_load(float* dst, const float* src, int n, int tid, int nthreads) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = tid; i < n4; i += nthreads)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
        for (int i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = src[i];
    } else {
        for (int i = tid; i < n; i += nthreads) dst[i] = src[i];
    }
}
__device__ __forceinline__ void tile_store(float* dst, const float* src, int n, int tid, int nthreads) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = tid; i < n4; i += nthreads)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
        for (int i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = src[i];
    } else {
        for (int i = tid; i < n; i += nthreads) dst[i] = src[i];
    }
}

constexpr int FPB = 64;     // frames per tile
constexpr int DT = 128;     // threads per CTA

// ------------------------------------------------------------------------------------------------
// Upper decode: o [F,87] -> q [F,14,3,3], l [F,15,3]   (MLPHead tail Net/Upper_Net.py:355-364, ForKinematics :122-144,
// Transform2R Utils.py:274-281)
// ------------------------------------------------------------------------------------------------
struct UpperDecodeSmem {
    float in[FPB * 87];
    float q[FPB * 126];
    float l[FPB * 45];
    float rt[FPB * 12];
};

__global__ void __launch_bounds__(DT) upper_decode_kernel(const float* __restrict__ o, const float* __restrict__ body,
                                                          const float* __restrict__ R, const float* __restrict__ t,
                                                          float* __restrict__ lout, float* __restrict__ qout,
                                                          long long F, int L, int mode, long long row_offset,
                                                          int B_global) {
    MMEGO_DYN_SMEM(UpperDecodeSmem, sp);
    UpperDecodeSmem& s = *sp;
    const int tid = threadIdx.x;
    const long long f0 = (long long)blockIdx.x * FPB;
    const int nf = (int)((F - f0) < FPB ? (F - f0) : FPB);
    tile_load(s.in, o + f0 * 87, nf * 87, tid, DT);
    for (int i = tid; i < nf * 9; i += DT) s.rt[(i / 9) * 12 + i % 9] = R[f0 * 9 + i];
    for (int i = tid; i < nf * 3; i += DT) s.rt[(i / 3) * 12 + 9 + i % 3] = t[f0 * 3 + i];
    __syncthreads();
    if (tid < nf) {
        const long long r = row_offset + f0 + tid;
        const long long bi = (mode == 0) ? (r % B_global) : (r / L);
        const float* bd = body + bi * 60;
        const float* in = s.in + tid * 87;
        // joints: row stride 45 (odd) -> conflict-free
        upper_frame_decode(in, bd, s.rt + tid * 12, s.l + tid * 45, s.q + tid * 126);
    }
    __syncthreads();
    tile_store(lout + f0 * 45, s.l, nf * 45, tid, DT);
    if (qout) tile_store(qout + f0 * 126, s.q, nf * 126, tid, DT);
}

// ------------------------------------------------------------------------------------------------
// Lower decode: o [F,42] -> q [F,6,3,3], l [F,8,3]   (FusionModule tail Net/Lower_Net.py:126-135, ForKinematics :12-37)
// ------------------------------------------------------------------------------------------------
struct LowerDecodeSmem {
    float in[FPB * 42];
    float q[FPB * 54];
    float l[FPB * 24];
    float rt[FPB * 12];
};

__global__ void __launch_bounds__(DT) lower_decode_kernel(const float* __restrict__ o, const float* __restrict__ body,
                                                          const float* __restrict__ R, const float* __restrict__ t,
                                                          float* __restrict__ lout, float* __restrict__ qout,
                                                          long long F, int L, int mode, long long row_offset,
                                                          int B_global) {
    MMEGO_DYN_SMEM(LowerDecodeSmem, sp);
    LowerDecodeSmem& s = *sp;
    const int tid = threadIdx.x;
    const long long f0 = (long long)blockIdx.x * FPB;
    const int nf = (int)((F - f0) < FPB ? (F - f0) : FPB);
    tile_load(s.in, o + f0 * 42, nf * 42, tid, DT);
    for (int i = tid; i < nf * 9; i += DT) s.rt[(i / 9) * 12 + i % 9] = R[f0 * 9 + i];
    for (int i = tid; i < nf * 3; i += DT) s.rt[(i / 3) * 12 + 9 + i % 3] = t[f0 * 3 + i];
    __syncthreads();
    if (tid < nf) {
        const long long r = row_offset + f0 + tid;
        const long long bi = (mode == 0) ? (r % B_global) : (r / L);
        const float* bd = body + bi * 60;
        const float* in = s.in + tid * 42;
        lower_frame_decode(in, bd, s.rt + tid * 12, s.l + tid * 24, s.q + tid * 54);
    }
    __syncthreads();
    tile_store(lout + f0 * 24, s.l, nf * 24, tid, DT);
    if (qout) tile_store(qout + f0 * 54, s.q, nf * 54, tid, DT);
}

// ------------------------------------------------------------------------------------------------
// Assembly + metrics (Processor/Test/Demo_test.py:121-123, 64-69, 150-158)
// ------------------------------------------------------------------------------------------------
struct MetricsSmem {
    float up[FPB * 45];
    float lo[FPB * 24];
    float tg[FPB * 63];
    float pr[FPB * 63];
    double acc[kSumsLen];
};

__global__ void __launch_bounds__(DT) assemble_metrics_kernel(const float* __restrict__ up, const float* __restrict__ lo,
                                                              const float* __restrict__ tg, float* __restrict__ pred,
                                                              double* __restrict__ sums, long long F) {
    MMEGO_DYN_SMEM(MetricsSmem, sp);
    MetricsSmem& s = *sp;
    const int tid = threadIdx.x;
    const long long f0 = (long long)blockIdx.x * FPB;
    const int nf = (int)((F - f0) < FPB ? (F - f0) : FPB);
    const bool do_metrics = tg != nullptr && sums != nullptr;
    tile_load(s.up, up + f0 * 45, nf * 45, tid, DT);
    tile_load(s.lo, lo + f0 * 24, nf * 24, tid, DT);
    if (do_metrics) tile_load(s.tg, tg + f0 * 63, nf * 63, tid, DT);
    if (tid < kSumsLen) s.acc[tid] = 0.0;
    __syncthreads();
    if (tid < nf) {
        const float* u = s.up + tid * 45;
        const float* l = s.lo + tid * 24;
        float* p = s.pr + tid * 63;
        assemble_frame(u, l, p);
    }
    __syncthreads();
    if (pred) tile_store(pred + f0 * 63, s.pr, nf * 63, tid, DT);
    if (!do_metrics) return;
    float vals[kSumsLen];
#pragma unroll
    for (int i = 0; i < kSumsLen; ++i) vals[i] = 0.f;
    if (tid < nf) {
        frame_metrics(s.pr + tid * 63, s.tg + tid * 63, s.up + tid * 45, s.lo + tid * 24, vals);
    }
#pragma unroll
    for (int i = 0; i < kSumsLen; ++i) {
        float v = vals[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s.acc[i], (double)v);
    }
    __syncthreads();
    if (tid < kSumsLen && s.acc[tid] != 0.0) atomicAdd(&sums[tid], s.acc[tid]);
}

// ------------------------------------------------------------------------------------------------
// IMU_Net tail in fp32: used by the small-batch latency path (lstm_resident.cu) and by the fp32 FFMA test generation
// ------------------------------------------------------------------------------------------------
// attention pooling over the n samples of a frame (Net/IMU_Net.py:82-83): y [F,n,1024] -> s [F,1024]
constexpr int PW = 256;
__global__ void __launch_bounds__(PW) imu_pool_kernel(const float* __restrict__ y, const float* __restrict__ attn,
                                                      float* __restrict__ out, long long F, int n) {
    __shared__ float sc[64];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long f = blockIdx.x;
    const float* yf = y + f * (long long)n * 1024;
    float4 aw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[i] = reinterpret_cast<const float4*>(attn)[i * 32 + lane];
    const float ab = attn[1024];
    for (int sidx = w; sidx < n; sidx += PW / 32) {
        const float4* yp = reinterpret_cast<const float4*>(yf + (long long)sidx * 1024);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = yp[i * 32 + lane];
            a = fmaf(v.x, aw[i].x, a); a = fmaf(v.y, aw[i].y, a); a = fmaf(v.z, aw[i].z, a); a = fmaf(v.w, aw[i].w, a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sc[sidx] = a + ab;
    }
    __syncthreads();
    float m = -INFINITY;
    for (int i = 0; i < n; ++i) m = fmaxf(m, sc[i]);
    float sum = 0.f;
    for (int i = 0; i < n; ++i) sum += expf(sc[i] - m);
    const float inv = 1.0f / sum;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n; ++i) {
        const float wgt = expf(sc[i] - m) * inv;
        const float4 v = reinterpret_cast<const float4*>(yf + (long long)i * 1024)[tid];
        acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
        acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
    }
    reinterpret_cast<float4*>(out + f * 1024)[tid] = acc;
}

// fc2 + ortho6d (Net/IMU_Net.py:87-93, 7-47): g [F,1024] -> R [F,3,3], t [F,3]; one warp per frame
__global__ void __launch_bounds__(256) imu_decode_kernel(const float* __restrict__ g, const float* __restrict__ fc2,
                                                         float* __restrict__ R, float* __restrict__ t, long long F) {
    const int lane = threadIdx.x & 31;
    const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (f >= F) return;   // whole warp exits together; no block barrier below
    float4 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = reinterpret_cast<const float4*>(g + f * 1024)[i * 32 + lane];
    float T9[9];
#pragma unroll
    for (int o = 0; o < 9; ++o) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 wv = reinterpret_cast<const float4*>(fc2 + o * 1024)[i * 32 + lane];
            a = fmaf(wv.x, x[i].x, a); a = fmaf(wv.y, x[i].y, a); a = fmaf(wv.z, x[i].z, a); a = fmaf(wv.w, x[i].w, a);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        T9[o] = a + fc2[9 * 1024 + o];
    }
    if (lane == 0) {
        float m[9];
        ortho6d(T9, 1e-8f, m);
#pragma unroll
        for (int k = 0; k < 9; ++k) R[f * 9 + k] = m[k];
        t[f * 3] = T9[6]; t[f * 3 + 1] = T9[7]; t[f * 3 + 2] = T9[8];
    }
}

// ------------------------------------------------------------------------------------------------
// rigid transforms (Utils.py:274-292) as standalone ops for the drop-in Util module
// ------------------------------------------------------------------------------------------------
__global__ void transform2h_kernel(float* __restrict__ pts, const float* __restrict__ R, const float* __restrict__ t,
                                   long long total, int n, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long f = i / n;
    const float* r = R + f * 9;
    const float* tt = t + f * 3;
    float* p = pts + i * D;
    const float dx = p[0] - tt[0], dy = p[1] - tt[1], dz = p[2] - tt[2];
    p[0] = r[0] * dx + r[1] * dy + r[2] * dz;
    p[1] = r[3] * dx + r[4] * dy + r[5] * dz;
    p[2] = r[6] * dx + r[7] * dy + r[8] * dz;
}
__global__ void transform2r_kernel(const float* __restrict__ pts, const float* __restrict__ R,
                                   const float* __restrict__ t, float* __restrict__ out, long long total, int n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long f = i / n;
    const float* r = R + f * 9;
    const float* tt = t + f * 3;
    const float* p = pts + i * 3;
    out[i * 3] = r[0] * p[0] + r[3] * p[1] + r[6] * p[2] + tt[0];
    out[i * 3 + 1] = r[1] * p[0] + r[4] * p[1] + r[7] * p[2] + tt[1];
    out[i * 3 + 2] = r[2] * p[0] + r[5] * p[1] + r[8] * p[2] + tt[2];
}

}  // namespace

void launch_upper_decode(const float* o, const float* body, const float* R, const float* t, float* l, float* q,
                         long long F, int L, int mode, long long row_offset, int B_global, cudaStream_t st) {
    if (F <= 0) return;
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set)) {
        cudaFuncSetAttribute(upper_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UpperDecodeSmem));
    }
    MMEGO_LAUNCH(upper_decode_kernel, dim3((unsigned)((F + FPB - 1) / FPB)), dim3(DT), sizeof(UpperDecodeSmem), st, o,
                 body, R, t, l, q, F, L, mode, row_offset, B_global);
}
void launch_lower_decode(const float* o, const float* body, const float* R, const float* t, float* l, float* q,
                         long long F, int L, int mode, long long row_offset, int B_global, cudaStream_t st) {
    if (F <= 0) return;
    MMEGO_LAUNCH(lower_decode_kernel, dim3((unsigned)((F + FPB - 1) / FPB)), dim3(DT), sizeof(LowerDecodeSmem), st, o,
                 body, R, t, l, q, F, L, mode, row_offset, B_global);
}
void launch_assemble_metrics(const float* up, const float* lo, const float* tg, float* pred, double* sums, long long F,
                             cudaStream_t st) {
    if (F <= 0) return;
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set)) {
        cudaFuncSetAttribute(assemble_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MetricsSmem));
    }
    MMEGO_LAUNCH(assemble_metrics_kernel, dim3((unsigned)((F + FPB - 1) / FPB)), dim3(DT), sizeof(MetricsSmem), st, up,
                 lo, tg, pred, sums, F);
}
void launch_imu_pool(const float* y, const float* attn, float* out, long long F, int n, cudaStream_t st) {
    if (F <= 0) return;
    MMEGO_LAUNCH(imu_pool_kernel, dim3((unsigned)F), dim3(PW), 0, st, y, attn, out, F, n);
}
void launch_imu_decode(const float* g, const float* fc2, float* R, float* t, long long F, cudaStream_t st) {
    if (F <= 0) return;
    MMEGO_LAUNCH(imu_decode_kernel, dim3((unsigned)((F + 7) / 8)), dim3(256), 0, st, g, fc2, R, t, F);
}
void launch_transform2h(float* pts, const float* R, const float* t, long long F, int n, int D, cudaStream_t st) {
    const long long total = F * n;
    if (total <= 0) return;
    MMEGO_LAUNCH(transform2h_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, pts, R, t, total, n, D);
}
void launch_transform2r(const float* pts, const float* R, const float* t, float* out, long long F, int n,
                        cudaStream_t st) {
    const long long total = F * n;
    if (total <= 0) return;
    MMEGO_LAUNCH(transform2r_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, pts, R, t, out, total, n);
}

}  // namespace mmego
