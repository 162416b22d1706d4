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
#include "mma_frag.cuh"
#ifndef MMEGO_EMUL
#include "tc_common.cuh"
#endif

namespace mmego {

namespace {

#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
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


#endif  // MMEGO_FFMA_GEN

// ================================================================================================================
// Tensor-core version (default): the same network on mma.sync m16n8k16 fragments (mma_frag.cuh), fp16 hi/lo split
// products with fp32 accumulation.  256 threads = 8 warps per frame chunk of 128 points; a warp owns a tile of 16
// points and carries it through all six layers in registers (the accumulator fragments of one layer are the A
// fragments of the next).  Softmax attention pooling is flash-style: every warp keeps (max, sum, 64 weighted channel
// sums) for its own points and the 8 warps are merged once per frame through shared memory.
// Per 16-point tile: 150 MMAs (50 fragment products x 3) against 6,144 FFMAs per POINT in the kernel above.
// ================================================================================================================
using UM = UpperMmaLayout;
constexpr int MT = 128;                 // threads per CTA
constexpr int MW = MT / 32;             // warps; a warp carries TWO 16-point tiles at a time (32 points)
constexpr int PART_LD = 68;             // [64 channel sums | max | sum | pad]

struct MmaSmem {
    uint32_t w[UM::TOTAL];              // fragment-ordered weights, biases, attention vector, scales
    float tile[MW][32][8];              // transformed input points of each warp's tile pair (6 channels + 2 zero)
    float part[2][MW][PART_LD];         // per-warp softmax partials, double-buffered by frame parity
    unsigned long long bar[2];          // mbarriers of the two cloud buffers
    // followed by float cloud[2][N * 6]: the frame's radar cloud, staged by the TMA unit (1-D bulk copy) one frame ahead
};

// Two tiles per warp: every B fragment read from shared memory feeds two MMA chains (half the fragment loads per MMA)
// and the warp always has two independent dependency chains in flight -- with one tile per warp the kernel sat at 47 %
// tensor-pipe / 53 % issue utilisation (profiles/r01small2_ncu_summary.txt): latency-bound on the layer-to-layer chain.
__global__ void __launch_bounds__(MT, 2) upper_point_mma_kernel(float* __restrict__ x, const float* __restrict__ R,
                                                                const float* __restrict__ t,
                                                                const float* __restrict__ wblob,
                                                                float* __restrict__ gout, float* __restrict__ gw,
                                                                long long F, int N, int stage_clouds) {
    MMEGO_DYN_SMEM(MmaSmem, sp);
    MmaSmem& s = *sp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    for (int i = tid * 4; i < UM::TOTAL; i += MT * 4)
        *reinterpret_cast<uint4*>(s.w + i) = *reinterpret_cast<const uint4*>(wblob + i);
    __syncthreads();
    const uint4* wf = reinterpret_cast<const uint4*>(s.w);
    const float* wfl = reinterpret_cast<const float*>(s.w);
    const float* osc = wfl + UM::OS;
    // Radar clouds are staged by TMA: while the warps work on frame f, one thread has the copy of the CTA's next frame
    // (N x 24 bytes, one bulk transfer) in flight into the other buffer; the per-lane global loads that used to open
    // every tile's dependency chain become shared-memory reads.  (Odd N breaks the 16-byte rule of bulk copies: the
    // lanes then read the cloud from global memory as before.)
    float* cloud = reinterpret_cast<float*>(sp + 1);
    const uint32_t cloud_bytes = (uint32_t)N * 24u;
#ifndef MMEGO_EMUL
    const bool staged = stage_clouds && (cloud_bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.bar);
    if (staged) {
        if (tid == 0) {
            tc::mbar_init(&bars[0], 1);
            tc::mbar_init(&bars[1], 1);
            tc::fence_barrier_init();
        }
        __syncthreads();
        if (tid == 0 && (long long)blockIdx.x < F) {
            tc::mbar_expect_tx(&bars[0], cloud_bytes);
            tc::tma_load_1d(cloud, x + (long long)blockIdx.x * N * 6, cloud_bytes, &bars[0]);
        }
    }
#else
    const bool staged = false;
#endif

    int parity = 0;
    for (long long f = blockIdx.x; f < F; f += gridDim.x, parity ^= 1) {
        float* xf = x + f * (long long)N * 6;
        const float* xin = xf;
#ifndef MMEGO_EMUL
        if (staged) {
            // buffer `parity` holds frame f; buffer parity^1 was last read two frames ago (a __syncthreads ends every frame)
            if (tid == 0 && f + gridDim.x < F) {
                tc::mbar_expect_tx(&bars[parity ^ 1], cloud_bytes);
                tc::tma_load_1d(cloud + (size_t)(parity ^ 1) * N * 6, x + (f + gridDim.x) * (long long)N * 6, cloud_bytes,
                                &bars[parity ^ 1]);
            }
            const long long it = (f - blockIdx.x) / gridDim.x;            // this CTA's frame counter: buffer parity = it & 1
            tc::mbar_wait(&bars[parity], (uint32_t)((it >> 1) & 1));
            xin = cloud + (size_t)parity * N * 6;
        }
#endif
        float* gwf = gw ? gw + f * (long long)N : nullptr;
        float run_m = -INFINITY, run_s = 0.f;
        float gp[16];                    // weighted sums of channels 8j + 2*tq + {0,1}, over this lane's rows
#pragma unroll
        for (int i = 0; i < 16; ++i) gp[i] = 0.f;
        float rt[12];
        {
            const float* Rf = R + f * 9;
            const float* tf = t + f * 3;
#pragma unroll
            for (int k = 0; k < 9; ++k) rt[k] = __ldg(Rf + k);
#pragma unroll
            for (int k = 0; k < 3; ++k) rt[9 + k] = __ldg(tf + k);
        }

        for (int p0 = 0; p0 < N; p0 += MW * 32) {
            const int pbase = p0 + warp * 32;
            if (pbase >= N) break;       // warp-uniform: no live point in this tile pair
            // ---- load + Transform2H (in place) of the pair's 32 points: one point per lane -----------------------
            __syncwarp();
            {
                const int p = pbase + lane;
                float in[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (p < N) {
                    const float2 v0 = *reinterpret_cast<const float2*>(xin + p * 6);
                    const float2 v1 = *reinterpret_cast<const float2*>(xin + p * 6 + 2);
                    const float2 v2 = *reinterpret_cast<const float2*>(xin + p * 6 + 4);
                    const float dx = v0.x - rt[9], dy = v0.y - rt[10], dz = v1.x - rt[11];
                    in[0] = rt[0] * dx + rt[1] * dy + rt[2] * dz;
                    in[1] = rt[3] * dx + rt[4] * dy + rt[5] * dz;
                    in[2] = rt[6] * dx + rt[7] * dy + rt[8] * dz;
                    in[3] = v1.y; in[4] = v2.x; in[5] = v2.y;
                    *reinterpret_cast<float2*>(xf + p * 6) = make_float2(in[0], in[1]);
                    xf[p * 6 + 2] = in[2];
                }
                *reinterpret_cast<float4*>(&s.tile[warp][lane][0]) = make_float4(in[0], in[1], in[2], in[3]);
                *reinterpret_cast<float4*>(&s.tile[warp][lane][4]) = make_float4(in[4], in[5], 0.f, 0.f);
            }
            __syncwarp();
            // ---- layer-1 A fragments (tile u = points pbase + 16u ..): rows g, g+8; channels 2tq, 2tq+1 ------------
            uint32_t a1h[2][1][4], a1l[2][1][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float2 r0 = *reinterpret_cast<const float2*>(&s.tile[warp][16 * u + g][2 * tq]);
                const float2 r1 = *reinterpret_cast<const float2*>(&s.tile[warp][16 * u + g + 8][2 * tq]);
                frag::split2(r0.x, r0.y, a1h[u][0][0], a1l[u][0][0]);
                frag::split2(r1.x, r1.y, a1h[u][0][1], a1l[u][0][1]);
                a1h[u][0][2] = a1h[u][0][3] = a1l[u][0][2] = a1l[u][0][3] = 0u;
            }
            float psi[2][UM::NT6][4];
            {
                float c1[2][UM::NT1][4];
                frag::dense_tile2<UM::KS1, UM::NT1, true>(wf + UM::F1 / 4, wfl + UM::BI1, osc[0], a1h, a1l, c1, lane);
                uint32_t a2h[2][UM::KS2][4], a2l[2][UM::KS2][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) frag::to_afrag<UM::NT1, UM::KS2>(c1[u], a2h[u], a2l[u]);
                float c2[2][UM::NT2][4];
                frag::dense_tile2<UM::KS2, UM::NT2, true>(wf + UM::F2 / 4, wfl + UM::BI2, osc[1], a2h, a2l, c2, lane);
                uint32_t a3h[2][UM::KS3][4], a3l[2][UM::KS3][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) frag::to_afrag<UM::NT2, UM::KS3>(c2[u], a3h[u], a3l[u]);
                float c3[2][UM::NT3][4];
                frag::dense_tile2<UM::KS3, UM::NT3, true>(wf + UM::F3 / 4, wfl + UM::BI3, osc[2], a3h, a3l, c3, lane);
                // layer 4 reads [feat 0..23 | x[0:4] | pad]: k-step 1 = feat 16..23 and, in its upper half, the
                // first four input channels, which are exactly words 0/1 of the layer-1 fragment of lanes tq < 2
                uint32_t a4h[2][UM::KS4][4], a4l[2][UM::KS4][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    frag::to_afrag<UM::NT3, UM::KS4>(c3[u], a4h[u], a4l[u]);
                    a4h[u][1][2] = tq < 2 ? a1h[u][0][0] : 0u;
                    a4h[u][1][3] = tq < 2 ? a1h[u][0][1] : 0u;
                    a4l[u][1][2] = tq < 2 ? a1l[u][0][0] : 0u;
                    a4l[u][1][3] = tq < 2 ? a1l[u][0][1] : 0u;
                }
                float c4[2][UM::NT4][4];
                frag::dense_tile2<UM::KS4, UM::NT4, true>(wf + UM::F4 / 4, wfl + UM::BI4, osc[3], a4h, a4l, c4, lane);
                uint32_t a5h[2][UM::KS5][4], a5l[2][UM::KS5][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) frag::to_afrag<UM::NT4, UM::KS5>(c4[u], a5h[u], a5l[u]);
                float c5[2][UM::NT5][4];
                frag::dense_tile2<UM::KS5, UM::NT5, true>(wf + UM::F5 / 4, wfl + UM::BI5, osc[4], a5h, a5l, c5, lane);
                uint32_t a6h[2][UM::KS6][4], a6l[2][UM::KS6][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) frag::to_afrag<UM::NT5, UM::KS6>(c5[u], a6h[u], a6l[u]);
                frag::dense_tile2<UM::KS6, UM::NT6, true>(wf + UM::F6 / 4, wfl + UM::BI6, osc[5], a6h, a6l, psi, lane);
            }
            // ---- attention scores of this lane's four rows (tile u: rows g, g+8) (fp32 FFMA; quad reduce) ----------
            float sc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
            for (int j = 0; j < UM::NT6; ++j) {
                const float2 wa = *reinterpret_cast<const float2*>(wfl + UM::WA + 8 * j + 2 * tq);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    sc[u][0] = fmaf(wa.x, psi[u][j][0], sc[u][0]);
                    sc[u][0] = fmaf(wa.y, psi[u][j][1], sc[u][0]);
                    sc[u][1] = fmaf(wa.x, psi[u][j][2], sc[u][1]);
                    sc[u][1] = fmaf(wa.y, psi[u][j][3], sc[u][1]);
                }
            }
            const float ba = wfl[UM::BA];
            bool live[2][2];
            float cm = -INFINITY;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    float v = sc[u][hrow];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int p = pbase + 16 * u + g + 8 * hrow;
                    live[u][hrow] = p < N;
                    v = live[u][hrow] ? v + ba : -INFINITY;
                    sc[u][hrow] = v;
                    if (gwf && tq == 0 && live[u][hrow]) gwf[p] = v;   // raw scores; normalised once max and sum are known
                    cm = fmaxf(cm, v);
                }
            // ---- online softmax over this warp's points -----------------------------------------------------------
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
            const float new_m = fmaxf(run_m, cm);          // finite: row 0 of the pair is live
            float e[2][2];
            float es = 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    e[u][hrow] = live[u][hrow] ? expf(sc[u][hrow] - new_m) : 0.f;
                    es += e[u][hrow];
                }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
            const float rescale = (run_m == -INFINITY) ? 0.f : expf(run_m - new_m);
#pragma unroll
            for (int j = 0; j < UM::NT6; ++j) {
                float a0 = fmaf(e[0][0], psi[0][j][0], e[0][1] * psi[0][j][2]);
                float a1 = fmaf(e[0][0], psi[0][j][1], e[0][1] * psi[0][j][3]);
                a0 = fmaf(e[1][0], psi[1][j][0], fmaf(e[1][1], psi[1][j][2], a0));
                a1 = fmaf(e[1][0], psi[1][j][1], fmaf(e[1][1], psi[1][j][3], a1));
                gp[2 * j] = fmaf(gp[2 * j], rescale, a0);
                gp[2 * j + 1] = fmaf(gp[2 * j + 1], rescale, a1);
            }
            run_s = fmaf(run_s, rescale, es);
            run_m = new_m;
        }
        // ---- merge the warps ------------------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float v = gp[i];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            gp[i] = v;
        }
        float* mine = s.part[parity][warp];
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < UM::NT6; ++j)
                *reinterpret_cast<float2*>(mine + 8 * j + 2 * tq) = make_float2(gp[2 * j], gp[2 * j + 1]);
            if (tq == 0) { mine[64] = run_m; mine[65] = run_s; }
        }
        __syncthreads();
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < MW; ++w) M = fmaxf(M, s.part[parity][w][64]);
        float S = 0.f, G = 0.f;
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            const float mw = s.part[parity][w][64];
            const float scw = (mw == -INFINITY) ? 0.f : expf(mw - M);
            S = fmaf(scw, s.part[parity][w][65], S);
            if (tid < 64) G = fmaf(scw, s.part[parity][w][tid], G);
        }
        const float inv = __fdividef(1.0f, S);      // S in [1, N]; no IEEE slow path between warp-wide MMAs (see lstm_small.cu)
        if (tid < 64) gout[f * 64 + tid] = G * inv;
        if (gwf)
            for (int p = tid; p < N; p += MT) gwf[p] = expf(gwf[p] - M) * inv;
    }
}

}  // namespace

#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
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

#endif  // MMEGO_FFMA_GEN

// wblob: UpperMmaLayout (pack_upper_point_mma)
void launch_upper_point_mma(float* x, const float* R, const float* t, const float* wblob, float* g, float* gw,
                            long long F, int N, int sm_count, int stage_clouds, cudaStream_t st) {
    if (F <= 0) return;
    static bool attr_set[64] = {false};
    const size_t smem = sizeof(MmaSmem) + 2 * (size_t)N * 6 * sizeof(float);       // + the two staged cloud buffers
    static int attr_bytes[64] = {0};
    int d = 0;
    cudaGetDevice(&d);
    if (attr_bytes[d & 63] < (int)smem) {
        cudaFuncSetAttribute(upper_point_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_bytes[d & 63] = (int)smem;
    }
    long long grid = F < (long long)sm_count * 2 ? F : (long long)sm_count * 2;     // 2 CTAs of 128 threads x 250 registers per SM
    MMEGO_LAUNCH(upper_point_mma_kernel, dim3((unsigned)grid), dim3(MT), smem, st, x, R, t, wblob, g, gw, F, N, stage_clouds);
}

}  // namespace mmego
