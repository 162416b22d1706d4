// gemm_ffma.cu -- fp32 FFMA tiled GEMM with multi-segment K and fused epilogues.
//
//   C[M,N] = sum_seg A_seg[M,K_seg] * W[N, Kp]^T            (W packed row-major, segments padded to 16)
//
// Used for: every Linear / 1x1 conv / (9,1) temporal conv (as 9 row-shifted segments + the residual 1x1 conv as
// a 10th) of the ST-GCN, the input projections of the H=64 LSTMs, the heads, and -- with the LSTM-cell epilogue --
// the exact-fp32 recurrent step of IMU_Net's H=512 LSTMs (the tcgen05 kernel in lstm_tc.cu is the fast path).
//
// Tile: 128 (M) x BN (N) x 16 (K), 256 threads, 8 x BN/16 accumulators per thread, double-buffered shared memory
// with register prefetch (one __syncthreads per K tile).  Column ownership is "4 groups of BN/4" so that with the
// gate-interleaved LSTM weight packing one thread holds the i,f,g,o pre-activations of the same hidden unit.
#include "internal.h"

#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
namespace mmego {

namespace {

constexpr int BM = 128, BK = 16, NT = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(NT) gemm_ffma_kernel(GemmBatch batch) {
    constexpr int TN = BN / 16;                 // accumulator columns per thread
    constexpr int CPG = (TN >= 4) ? TN / 4 : 1; // columns per group per thread
    constexpr int NG = TN / CPG;                // groups (4, or 2 for BN=32)
    constexpr int GW = BN / NG;                 // group width in columns
    constexpr int WPT = BN * BK / NT;           // W floats loaded per thread per tile (8, 4, 2)
    const GemmArgs& g = batch.g[blockIdx.z];

    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // loader mapping
    const int a_row = tid % BM, a_k = (tid / BM) * 8;
    const int w_n = tid % BN, w_k = (tid / BN) * WPT;
    const long long grow = (long long)m0 + a_row;
    const bool row_ok = grow < g.M;
    const bool wn_ok = (n0 + w_n) < g.N;
    // compute mapping
    const int ty = tid / 16, tx = tid % 16;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[8], rw[WPT];
    int seg = 0, kin = 0, kglob = 0;   // current segment, k offset inside it, k offset in W

    auto load_tile = [&](int s, int kk, int kg) {
        const GemmSeg& sg = g.seg[s];
        bool ok = row_ok;
        long long src = grow + sg.shift;
        if (sg.period > 0) {
            int pos = (int)(grow % sg.period) + sg.shift;
            ok = ok && pos >= 0 && pos < sg.period;
        }
        const float* ap = sg.a + src * sg.lda + kk + a_k;
        if (ok && sg.vec && (kk + a_k + 8) <= sg.k) {
            float4 v0 = *reinterpret_cast<const float4*>(ap);
            float4 v1 = *reinterpret_cast<const float4*>(ap + 4);
            ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
            ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ra[i] = (ok && (kk + a_k + i) < sg.k) ? ap[i] : 0.f;
        }
        const float* wp = g.w + (long long)(n0 + w_n) * g.ldw + kg + w_k;
        if (wn_ok) {
            if constexpr (WPT == 8) {
                float4 v0 = *reinterpret_cast<const float4*>(wp);
                float4 v1 = *reinterpret_cast<const float4*>(wp + 4);
                rw[0] = v0.x; rw[1] = v0.y; rw[2] = v0.z; rw[3] = v0.w;
                rw[4] = v1.x; rw[5] = v1.y; rw[6] = v1.z; rw[7] = v1.w;
            } else if constexpr (WPT == 4) {
                float4 v0 = *reinterpret_cast<const float4*>(wp);
                rw[0] = v0.x; rw[1] = v0.y; rw[2] = v0.z; rw[3] = v0.w;
            } else {
                float2 v0 = *reinterpret_cast<const float2*>(wp);
                rw[0] = v0.x; rw[1] = v0.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < WPT; ++i) rw[i] = 0.f;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][a_k + i][a_row] = ra[i];
#pragma unroll
        for (int i = 0; i < WPT; ++i) Bs[buf][w_k + i][w_n] = rw[i];
    };
    auto advance = [&]() {   // move (seg, kin, kglob) to the next K tile; returns false at the end
        kin += BK;
        kglob += BK;
        if (kin >= g.seg[seg].kpad) { seg++; kin = 0; }
        return seg < g.nseg;
    };

    const int total_tiles = g.ktot / BK;
    load_tile(seg, kin, kglob);
    store_tile(0);
    __syncthreads();
    for (int t = 0; t < total_tiles; ++t) {
        const int buf = t & 1;
        bool more = advance();
        if (more) load_tile(seg, kin, kglob);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int gi = 0; gi < NG; ++gi) {
                if constexpr (CPG == 2) {
                    float2 v = *reinterpret_cast<const float2*>(&Bs[buf][k][gi * GW + tx * 2]);
                    b[gi * 2] = v.x;
                    b[gi * 2 + 1] = v.y;
                } else {
                    b[gi] = Bs[buf][k][gi * GW + tx];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store_tile(buf ^ 1);
        __syncthreads();
    }

    // ------------------------------------------------------------------------------------------ epilogue
    if constexpr (EPI == EPI_LSTM) {
        // column j = gate*CPG + e ; hidden unit u = tile*GW + tx*CPG + e ; packed bias index = n0 + gate*GW + tx*CPG + e
#pragma unroll
        for (int e = 0; e < CPG; ++e) {
            const int u = blockIdx.x * GW + tx * CPG + e;
            const int H = g.N / 4;
            if (u >= H) continue;
            const float bi = g.bias[n0 + 0 * GW + tx * CPG + e];
            const float bf = g.bias[n0 + 1 * GW + tx * CPG + e];
            const float bg = g.bias[n0 + 2 * GW + tx * CPG + e];
            const float bo = g.bias[n0 + 3 * GW + tx * CPG + e];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long row = (long long)m0 + ty * 8 + i;
                if (row >= g.M) continue;
                const float ig = sigmoidf_(acc[i][0 * CPG + e] + bi);
                const float fg = sigmoidf_(acc[i][1 * CPG + e] + bf);
                const float gg = tanhf(acc[i][2 * CPG + e] + bg);
                const float og = sigmoidf_(acc[i][3 * CPG + e] + bo);
                float* cs = g.cstate + row * g.ldcs + u;
                const float cprev = g.has_state ? *cs : 0.f;
                const float cn = fg * cprev + ig * gg;
                *cs = cn;
                g.c[row * g.ldc + u] = og * tanhf(cn);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long row = (long long)m0 + ty * 8 + i;
            if (row >= g.M) continue;
            const float* brow = g.bias + (g.rowmod > 0 ? (long long)(row % g.rowmod) * g.N : 0);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int col = n0 + (j / CPG) * GW + tx * CPG + (j % CPG);
                if (col >= g.N) continue;
                float v = acc[i][j] + brow[col];
                if (g.relu) v = fmaxf(v, 0.f);
                if (EPI == EPI_F6) {
                    const long long b = row / g.f6_period, pos = row % g.f6_period;
                    g.c[b * (long long)g.N * g.f6_period + (long long)col * g.f6_period + pos] = v;
                } else {
                    g.c[row * g.ldc + col] = v;
                }
            }
        }
    }
}

template <int BN, int EPI>
void launch_t(const GemmBatch& b, int nz, cudaStream_t st) {
    const GemmArgs& g = b.g[0];
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, nz);
    auto kfn = gemm_ffma_kernel<BN, EPI>;
    MMEGO_LAUNCH(kfn, grid, dim3(NT), 0, st, b);
}

}  // namespace

void launch_gemm(const GemmBatch& b, int nz, int bn, int epi, cudaStream_t st) {
    if (b.g[0].M <= 0 || b.g[0].N <= 0) return;
    if (epi == EPI_LSTM) {
        if (bn == 128) launch_t<128, EPI_LSTM>(b, nz, st);
        else launch_t<64, EPI_LSTM>(b, nz, st);
    } else if (epi == EPI_F6) {
        launch_t<64, EPI_F6>(b, nz, st);
    } else {
        if (bn == 128) launch_t<128, EPI_STORE>(b, nz, st);
        else if (bn == 64) launch_t<64, EPI_STORE>(b, nz, st);
        else launch_t<32, EPI_STORE>(b, nz, st);
    }
}

}  // namespace mmego
#endif  // MMEGO_FFMA_GEN
