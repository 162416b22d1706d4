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
#include "mma_frag.cuh"

namespace mmego {

namespace {

using LL = LowerFrameLayout;
constexpr int NT = 128;
// BasePointNet's folded weights (W1..B3, 2.8k floats) in constant memory: immediate FFMA operands, as in point_upper.cu.
// The three 64x64 projections stay in shared memory (they are indexed per thread).
constexpr int NMAX = 512;      // max points per frame supported by the rank select
constexpr int LDP = 65;
#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
constexpr int kMlpFloats = LL::WQ;
__constant__ float c_mlp[kMlpFloats];

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
            s.key[p] = nx == nx ? nx : INFINITY;     // NaN sorts first, as in torch.sort(descending): keeps the ranks a permutation
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


#endif  // MMEGO_FFMA_GEN

// ================================================================================================================
// Tensor-core version (default): the same per-frame work on mma.sync m16n8k16 fragments (mma_frag.cuh), fp16 hi/lo
// split products with fp32 accumulation.  128 threads = 4 warps per frame:
//   A  all threads: stage the frame's joint features, second in-place Transform2H, keys
//   B  all threads: rank select of the top 64;  then warps 0,1: to_k, warps 2,3: to_v of the 15 (16) joints
//   C  warp w: selected points 16w..16w+15 through BasePointNet -> P' = [feat61 | xyz] -> to_q -> q k^T -> softmax;
//      column sums of P and of alpha leave through quad shuffles
//   D  merge the four warps: a = sum P, a_T = (sum alpha) @ v, kbar
// Per frame: 4 x 186 + 192 = 936 MMAs instead of 0.61 M FFMAs.
// ================================================================================================================
using LM = LowerMmaLayout;
constexpr int KF_LD = 72;               // joint-feature row stride (floats): conflict-free float2 fragment reads
constexpr int KP_LD = 36;               // to_k(K) row stride in half2 words (32 + 4 pad): conflict-free B fragments
constexpr int kSmemWeightWords = LM::FK;   // layers 1-3 + to_q live in shared memory; to_k / to_v are read through L1

struct MmaCarve {
    uint32_t* w;
    float *pts, *key, *Kf, *vp, *psum, *asum;
    int* sel;
    uint32_t *kph, *kpl;
    unsigned long long* srt;   // composite sort keys, only carved for N > kRankSelectMax
    unsigned long long* xch;   // two 128-entry exchange buffers of the register sort, only carved for N <= kRegSortMax
    int P;                     // N rounded up to a power of two
    size_t bytes;
};
// Up to this many points the O(N^2) rank select (one pass, one barrier) is the cheaper top-64; above it the frame's keys
// are sorted as 64-bit composites (key descending, slot ascending) by a shared-memory bitonic network, O(N log^2 N).
constexpr int kRankSelectMax = 256;
// Up to one point per thread (the configuration's N = 128) the (key, slot) composites are sorted in REGISTERS: a 128-wide
// bitonic network whose 25 in-warp stages are shuffles and whose 3 cross-warp stages go through shared memory.  ~0.9k
// warp instructions per frame instead of the rank select's ~3.5k (it was 40 % of this kernel's issue slots).
constexpr int kRegSortMax = 128;
static_assert(kRegSortMax == 128, "sort128_top64 assumes 128 threads = 128 elements");

__host__ __device__ inline MmaCarve lower_mma_carve(unsigned char* base, int N) {
    MmaCarve c;
    size_t off = 0;
    auto take = [&](size_t n) { unsigned char* p = base ? base + off : nullptr; off += (n + 15) & ~size_t(15); return p; };
    c.w = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * kSmemWeightWords));
    c.Kf = reinterpret_cast<float*>(take(sizeof(float) * 16 * KF_LD));
    c.vp = reinterpret_cast<float*>(take(sizeof(float) * 16 * 64));
    c.kph = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * 16 * KP_LD));
    c.kpl = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * 16 * KP_LD));
    c.psum = reinterpret_cast<float*>(take(sizeof(float) * 4 * 64));
    c.asum = reinterpret_cast<float*>(take(sizeof(float) * 4 * 16));
    c.sel = reinterpret_cast<int*>(take(sizeof(int) * kLowerPts));
    c.key = reinterpret_cast<float*>(take(sizeof(float) * N));
    c.pts = reinterpret_cast<float*>(take(sizeof(float) * N * 6));
    c.P = 1;
    while (c.P < N) c.P <<= 1;
    c.srt = N > kRankSelectMax ? reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * c.P)) : nullptr;
    c.xch = N <= kRegSortMax ? reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * 2 * kRegSortMax)) : nullptr;
    c.bytes = off;
    return c;
}

// Top-64 of a frame with more than kRankSelectMax points.  Same total order as the rank select: key descending, equal
// keys by ascending slot (-0 == +0; NaN keys were replaced by +inf).  The float is mapped to an unsigned that ascends
// with it, inverted, and the slot goes in the low word; padding up to the power of two sorts last.  Not inlined: the
// main loop's register allocation stays what it is for the N <= 256 case.
__device__ __noinline__ void bitonic_top64(unsigned long long* srt, const float* key, int* sel, int N, int P) {
    const int tid = threadIdx.x;
    for (int i = tid; i < P; i += NT) {
        unsigned long long c = ~0ull;
        if (i < N) {
            const float kx = key[i];
            const uint32_t b = __float_as_uint(kx == 0.f ? 0.f : kx);
            const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
            c = ((unsigned long long)(~asc) << 32) | (uint32_t)i;
        }
        srt[i] = c;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < (P >> 1); i += NT) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const unsigned long long a = srt[lo], b = srt[hi];
                if ((a > b) == ((lo & k) == 0)) {
                    srt[lo] = b;
                    srt[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < kLowerPts; i += NT) sel[i] = (int)(uint32_t)srt[i];
}

// One element per thread, ascending by composite = (inverted sortable key, slot): position i ends up in thread i.
__device__ __forceinline__ void sort128_top64(const float* key, int* sel, unsigned long long* xch, int N) {
    const int i = threadIdx.x;
    uint32_t hi = 0xffffffffu, lo = (uint32_t)i;          // padding sorts last
    if (i < N) {
        const float kx = key[i];
        const uint32_t b = __float_as_uint(kx == 0.f ? 0.f : kx);
        hi = ~((b & 0x80000000u) ? ~b : (b | 0x80000000u));
    }
    int buf = 0;
#pragma unroll
    for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint32_t ph, pl;
            if (j < 32) {
                ph = __shfl_xor_sync(0xffffffffu, hi, j);
                pl = __shfl_xor_sync(0xffffffffu, lo, j);
            } else {
                unsigned long long* x = xch + buf * 128;
                buf ^= 1;
                x[i] = ((unsigned long long)hi << 32) | lo;
                __syncthreads();
                const unsigned long long p = x[i ^ j];
                ph = (uint32_t)(p >> 32);
                pl = (uint32_t)p;
            }
            const bool partner_less = ph < hi || (ph == hi && pl < lo);
            const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
            if (partner_less == keep_min) {
                hi = ph;
                lo = pl;
            }
        }
    }
    if (i < kLowerPts) sel[i] = (int)lo;
}

__global__ void __launch_bounds__(NT, 4) lower_frame_mma_kernel(float* __restrict__ x, const float* __restrict__ R,
                                                                const float* __restrict__ t,
                                                                const float* __restrict__ kfeat,
                                                                const float* __restrict__ wblob,
                                                                float* __restrict__ ak, long long F, int N) {
    MMEGO_DYN_SMEM(unsigned char, raw);
    const MmaCarve s = lower_mma_carve(raw, N);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    for (int i = tid * 4; i < kSmemWeightWords; i += NT * 4)
        *reinterpret_cast<uint4*>(s.w + i) = *reinterpret_cast<const uint4*>(wblob + i);
    for (int i = tid; i < KF_LD; i += NT) s.Kf[15 * KF_LD + i] = 0.f;     // joint 15 is padding
    const uint4* wf = reinterpret_cast<const uint4*>(s.w);
    const float* wfl = reinterpret_cast<const float*>(s.w);
    const uint4* wg = reinterpret_cast<const uint4*>(wblob);              // to_k / to_v fragments (global, L1-resident)
    const float os1 = wblob[LM::OS + 0], os2 = wblob[LM::OS + 1], os3 = wblob[LM::OS + 2], osq = wblob[LM::OS + 3];
    const float oskv = wblob[LM::OS + (warp < 2 ? 4 : 5)];

    for (long long f = blockIdx.x; f < F; f += gridDim.x) {
        __syncthreads();   // previous frame fully consumed (and weights staged on the first pass)
        // ---- A: joint features, second transform (in place), keys ------------------------------------------------
        for (int i = tid; i < kGcnV * 64; i += NT) s.Kf[(i >> 6) * KF_LD + (i & 63)] = kfeat[f * (kGcnV * 64) + i];
        {
            const float* Rf = R + f * 9;
            const float* tf = t + f * 3;
            const float r0 = __ldg(Rf), r1 = __ldg(Rf + 1), r2 = __ldg(Rf + 2), r3 = __ldg(Rf + 3), r4 = __ldg(Rf + 4),
                        r5 = __ldg(Rf + 5), r6 = __ldg(Rf + 6), r7 = __ldg(Rf + 7), r8 = __ldg(Rf + 8);
            const float t0 = __ldg(tf), t1 = __ldg(tf + 1), t2 = __ldg(tf + 2);
            float* xf = x + f * (long long)N * 6;
            for (int p = tid; p < N; p += NT) {
                const float2 v0 = *reinterpret_cast<const float2*>(xf + p * 6);
                const float2 v1 = *reinterpret_cast<const float2*>(xf + p * 6 + 2);
                const float2 v2 = *reinterpret_cast<const float2*>(xf + p * 6 + 4);
                const float dx = v0.x - t0, dy = v0.y - t1, dz = v1.x - t2;
                const float nx = r0 * dx + r1 * dy + r2 * dz;
                const float ny = r3 * dx + r4 * dy + r5 * dz;
                const float nz = r6 * dx + r7 * dy + r8 * dz;
                *reinterpret_cast<float2*>(xf + p * 6) = make_float2(nx, ny);
                xf[p * 6 + 2] = nz;
                float* pp = s.pts + p * 6;
                *reinterpret_cast<float2*>(pp) = make_float2(nx, ny);
                *reinterpret_cast<float2*>(pp + 2) = make_float2(nz, v1.y);
                *reinterpret_cast<float2*>(pp + 4) = v2;
                s.key[p] = nx == nx ? nx : INFINITY;     // NaN sorts first, as in torch.sort(descending): keeps the ranks a permutation
            }
        }
        __syncthreads();
        // ---- B1: rank select (equal keys: the lower slot wins) ---------------------------------------------------
        // rank(p) = #{q < p: key[q] >= key[p]} + #{q > p: key[q] > key[p]}; keys are read four at a time (broadcast
        // 16-byte loads), one compare per key: with 'x >= k' for slots below p and 'x > k' above, the tie rule needs no
        // second compare.
        if (N <= kRegSortMax) {
            sort128_top64(s.key, s.sel, s.xch, N);
        } else if (N <= kRankSelectMax) {
            for (int p = tid; p < N; p += NT) {
                const float kx = s.key[p];
                int rank = 0;
                const int n4 = N & ~3;
                for (int q = 0; q < n4; q += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(s.key + q);
                    if (q + 3 < p) {
                        rank += (k4.x >= kx) + (k4.y >= kx) + (k4.z >= kx) + (k4.w >= kx);
                    } else if (q > p) {
                        rank += (k4.x > kx) + (k4.y > kx) + (k4.z > kx) + (k4.w > kx);
                    } else {                                  // the group that contains p
                        rank += (q < p ? k4.x >= kx : (q > p && k4.x > kx));
                        rank += (q + 1 < p ? k4.y >= kx : (q + 1 > p && k4.y > kx));
                        rank += (q + 2 < p ? k4.z >= kx : (q + 2 > p && k4.z > kx));
                        rank += (q + 3 < p ? k4.w >= kx : (q + 3 > p && k4.w > kx));
                    }
                }
                for (int q = n4; q < N; ++q) {
                    const float kq = s.key[q];
                    rank += (q < p ? kq >= kx : (q > p && kq > kx));
                }
                if (rank < kLowerPts) s.sel[rank] = p;
            }
        } else {
            bitonic_top64(s.srt, s.key, s.sel, N, s.P);
        }
        // ---- B2: to_k (warps 0,1) / to_v (warps 2,3) of the joint features: 4 n-tiles each ----------------------
        {
            uint32_t kh[LM::KSP][4], kl[LM::KSP][4];
#pragma unroll
            for (int ks = 0; ks < LM::KSP; ++ks) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const float2 r0 = *reinterpret_cast<const float2*>(s.Kf + g * KF_LD + 16 * ks + 8 * hh + 2 * tq);
                    const float2 r1 = *reinterpret_cast<const float2*>(s.Kf + (g + 8) * KF_LD + 16 * ks + 8 * hh + 2 * tq);
                    frag::split2(r0.x, r0.y, kh[ks][2 * hh], kl[ks][2 * hh]);
                    frag::split2(r1.x, r1.y, kh[ks][2 * hh + 1], kl[ks][2 * hh + 1]);
                }
            }
            const bool isv = warp >= 2;
            const int j0 = (warp & 1) * 4;
            float o[4][4];
            frag::dense_tile<LM::KSP, 4, false, LM::NTP>(wg + (isv ? LM::FV : LM::FK) / 4, wblob + (isv ? LM::BV : LM::BK),
                                                         oskv, kh, kl, o, lane, j0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = 8 * (j0 + j) + 2 * tq;
                if (isv) {
                    *reinterpret_cast<float2*>(s.vp + g * 64 + col) = make_float2(o[j][0], o[j][1]);
                    *reinterpret_cast<float2*>(s.vp + (g + 8) * 64 + col) = make_float2(o[j][2], o[j][3]);
                } else {
                    uint32_t hi, lo;
                    frag::split2(o[j][0], o[j][1], hi, lo);
                    s.kph[g * KP_LD + (col >> 1)] = hi;
                    s.kpl[g * KP_LD + (col >> 1)] = lo;
                    frag::split2(o[j][2], o[j][3], hi, lo);
                    s.kph[(g + 8) * KP_LD + (col >> 1)] = hi;
                    s.kpl[(g + 8) * KP_LD + (col >> 1)] = lo;
                }
            }
        }
        __syncthreads();
        // ---- C: this warp's 16 selected points ---------------------------------------------------------------------
        {
            const float* pr0 = s.pts + s.sel[warp * 16 + g] * 6;
            const float* pr1 = s.pts + s.sel[warp * 16 + g + 8] * 6;
            uint32_t a1h[1][4], a1l[1][4];
            {
                float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
                if (tq < 3) {
                    v0 = *reinterpret_cast<const float2*>(pr0 + 2 * tq);
                    v1 = *reinterpret_cast<const float2*>(pr1 + 2 * tq);
                }
                frag::split2(v0.x, v0.y, a1h[0][0], a1l[0][0]);
                frag::split2(v1.x, v1.y, a1h[0][1], a1l[0][1]);
                a1h[0][2] = a1h[0][3] = a1l[0][2] = a1l[0][3] = 0u;
            }
            float P[LM::NT3][4];
            {
                float c1[LM::NT1][4];
                frag::dense_tile<LM::KS1, LM::NT1, true>(wf + LM::F1 / 4, wfl + LM::BI1, os1, a1h, a1l, c1, lane);
                uint32_t a2h[LM::KS2][4], a2l[LM::KS2][4];
                frag::to_afrag<LM::NT1, LM::KS2>(c1, a2h, a2l);
                float c2[LM::NT2][4];
                frag::dense_tile<LM::KS2, LM::NT2, true>(wf + LM::F2 / 4, wfl + LM::BI2, os2, a2h, a2l, c2, lane);
                uint32_t a3h[LM::KS3][4], a3l[LM::KS3][4];
                frag::to_afrag<LM::NT2, LM::KS3>(c2, a3h, a3l);
                frag::dense_tile<LM::KS3, LM::NT3, true>(wf + LM::F3 / 4, wfl + LM::BI3, os3, a3h, a3l, P, lane);
            }
            // P' columns 61..63 (zero so far: padded rows of layer 3) take x, y, z
            if (tq == 2) { P[7][1] = pr0[0]; P[7][3] = pr1[0]; }
            if (tq == 3) { P[7][0] = pr0[1]; P[7][1] = pr0[2]; P[7][2] = pr1[1]; P[7][3] = pr1[2]; }
            // column sums of P over the tile's 16 points -> psum[warp][channel in the reference's order xyz | feat]
#pragma unroll
            for (int j = 0; j < LM::NT3; ++j) {
                float v0 = P[j][0] + P[j][2], v1 = P[j][1] + P[j][3];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                }
                if (g == 0) {
                    const int c0 = 8 * j + 2 * tq, c1 = c0 + 1;
                    s.psum[warp * 64 + (c0 < 61 ? c0 + 3 : c0 - 61)] = v0;
                    s.psum[warp * 64 + (c1 < 61 ? c1 + 3 : c1 - 61)] = v1;
                }
            }
            // to_q, then scores against to_k(K)
            uint32_t qh[LM::KSP][4], ql[LM::KSP][4];
            {
                uint32_t ph[LM::KSP][4], pl[LM::KSP][4];
                frag::to_afrag<LM::NT3, LM::KSP>(P, ph, pl);
                float q[LM::NTP][4];
                frag::dense_tile<LM::KSP, LM::NTP, false>(wf + LM::FQ / 4, wfl + LM::BQ, osq, ph, pl, q, lane);
                frag::to_afrag<LM::NTP, LM::KSP>(q, qh, ql);
            }
            float sc[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float big[4] = {0.f, 0.f, 0.f, 0.f}, small[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < LM::KSP; ++ks) {
                    const int wi = (8 * j + g) * KP_LD + 8 * ks + tq;
                    const uint4 b = make_uint4(s.kph[wi], s.kph[wi + 4], s.kpl[wi], s.kpl[wi + 4]);
                    frag::mma3(big, small, qh[ks], ql[ks], b);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) sc[j][i] = (big[i] + small[i]) * 0.125f;      // dim_head^-0.5, Lower_Net.py:106
            }
            if (tq == 3) { sc[1][1] = -INFINITY; sc[1][3] = -INFINITY; }                   // joint 15 does not exist
            float al[2][2];                                                                // alpha summed over rows g, g+8
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float m = fmaxf(fmaxf(sc[0][2 * r], sc[0][2 * r + 1]), fmaxf(sc[1][2 * r], sc[1][2 * r + 1]));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                float e[4] = {expf(sc[0][2 * r] - m), expf(sc[0][2 * r + 1] - m), expf(sc[1][2 * r] - m),
                              expf(sc[1][2 * r + 1] - m)};
                float sum = (e[0] + e[1]) + (e[2] + e[3]);
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                const float inv = __fdividef(1.0f, sum);     // branch-free (no IEEE slow path between warp-wide MMAs)
                if (r == 0) { al[0][0] = e[0] * inv; al[0][1] = e[1] * inv; al[1][0] = e[2] * inv; al[1][1] = e[3] * inv; }
                else { al[0][0] += e[0] * inv; al[0][1] += e[1] * inv; al[1][0] += e[2] * inv; al[1][1] += e[3] * inv; }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float v = al[j][e];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (g == 0) s.asum[warp * 16 + 8 * j + 2 * tq + e] = v;
                }
        }
        __syncthreads();
        // ---- D: merge ---------------------------------------------------------------------------------------------------
        if (tid < 64) {
            ak[f * 192 + tid] = (s.psum[tid] + s.psum[64 + tid]) + (s.psum[128 + tid] + s.psum[192 + tid]);
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < kGcnV; ++j)
                a = fmaf((s.asum[j] + s.asum[16 + j]) + (s.asum[32 + j] + s.asum[48 + j]), s.vp[j * 64 + tid], a);
            ak[f * 192 + 64 + tid] = a;
        } else {
            const int c = tid - 64;
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < kGcnV; ++j) a += s.Kf[j * KF_LD + c];
            ak[f * 192 + 128 + c] = a * (1.0f / 15.0f);
        }
    }
}

}  // namespace

int lower_frame_max_points() { return NMAX; }

#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
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

#endif  // MMEGO_FFMA_GEN

// wblob: LowerMmaLayout (pack_lower_frame_mma)
void launch_lower_frame_mma(float* x, const float* R, const float* t, const float* kfeat, const float* wblob, float* ak,
                            long long F, int N, int sm_count, cudaStream_t st) {
    if (F <= 0) return;
    const size_t smem = lower_mma_carve(nullptr, N).bytes;
    static int attr_bytes[64] = {0};
    int d = 0;
    cudaGetDevice(&d);
    if (attr_bytes[d & 63] < (int)smem) {
        cudaFuncSetAttribute(lower_frame_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_bytes[d & 63] = (int)smem;
    }
    long long grid = F < (long long)sm_count * 4 ? F : (long long)sm_count * 4;
    MMEGO_LAUNCH(lower_frame_mma_kernel, dim3((unsigned)grid), dim3(NT), smem, st, x, R, t, kfeat, wblob, ak, F, N);
}

}  // namespace mmego
