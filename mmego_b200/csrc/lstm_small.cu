// lstm_small.cu -- persistent recurrent kernel for the H=64 bidirectional LSTMs (Upper_Net grnn, Net/Upper_Net.py:333;
// Lower_Net rnn_pk, Net/Lower_Net.py:91).  Kernel "K2-small".
//
// The input projection (x W_ih^T + b_ih + b_hh for all timesteps, both directions) is one batched GEMM
// (gemm_ffma.cu) whose output columns are already in this kernel's thread order.  This kernel then runs all T steps
// of one direction for a group of SEQ sequences without leaving the SM:
//   * thread (warp w, lane = gate*8 + e) owns gate row (gate, unit w*8+e) and keeps its 64 recurrent weights
//     W_hh[row, :] in REGISTERS for the whole sequence,
//   * h_{t-1} of the group lives in shared memory (double buffered, 128-bit broadcast reads),
//   * the four gates of a unit sit in one warp (lanes e, e+8, e+16, e+24) and meet through warp shuffles,
//   * c stays in the registers of lanes 0..7 of each warp.
// One __syncthreads per timestep.
#include "internal.h"

namespace mmego {

namespace {

constexpr int H = kSmallH;     // 64
constexpr int NTH = 256;       // 4 gates x 64 units
constexpr int SEQ = 8;         // sequences per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// gx   [S][T][2][256]   input projection incl. biases, column = thread order (w*32 + gate*8 + e)
// whh  [2][256][64]     recurrent weights, row = thread order
// h0/c0 [2][S][64] (this layer's slice of the [6,S,64] state) or nullptr for zeros
// y    [S][T][128]      fwd -> cols 0..63, bwd -> cols 64..127
// hn/cn [2][S][64] or nullptr
__global__ void __launch_bounds__(NTH) lstm_small_kernel(const float* __restrict__ gx, const float* __restrict__ whh,
                                                         const float* __restrict__ h0, const float* __restrict__ c0,
                                                         float* __restrict__ y, float* __restrict__ hn,
                                                         float* __restrict__ cn, int S, int T) {
    __shared__ __align__(16) float hs[2][SEQ][H];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gate = lane >> 3, e = lane & 7, unit = w * 8 + e;
    const int dir = blockIdx.y;
    const int s0 = blockIdx.x * SEQ;

    float wr[H];
    {
        const float* wp = whh + ((long long)dir * NTH + tid) * H;
#pragma unroll
        for (int k = 0; k < H; k += 4) {
            const float4 v = *reinterpret_cast<const float4*>(wp + k);
            wr[k] = v.x; wr[k + 1] = v.y; wr[k + 2] = v.z; wr[k + 3] = v.w;
        }
    }
    float c[SEQ];
#pragma unroll
    for (int s = 0; s < SEQ; ++s) {
        c[s] = 0.f;
        if (gate == 0 && (s0 + s) < S) {
            c[s] = c0 ? c0[((long long)dir * S + s0 + s) * H + unit] : 0.f;
            hs[0][s][unit] = h0 ? h0[((long long)dir * S + s0 + s) * H + unit] : 0.f;
        } else if (gate == 0) {
            hs[0][s][unit] = 0.f;
        }
    }
    __syncthreads();

    float gnext[SEQ];
    auto load_gx = [&](int step) {
        const int tt = dir ? (T - 1 - step) : step;
#pragma unroll
        for (int s = 0; s < SEQ; ++s)
            gnext[s] = (s0 + s) < S ? gx[(((long long)(s0 + s) * T + tt) * 2 + dir) * NTH + tid] : 0.f;
    };
    load_gx(0);
    for (int step = 0; step < T; ++step) {
        const int cur = step & 1;
        const int tt = dir ? (T - 1 - step) : step;
        float acc[SEQ];
#pragma unroll
        for (int s = 0; s < SEQ; ++s) acc[s] = gnext[s];
        if (step + 1 < T) load_gx(step + 1);
#pragma unroll
        for (int k = 0; k < H; k += 4) {
#pragma unroll
            for (int s = 0; s < SEQ; ++s) {
                const float4 hv = *reinterpret_cast<const float4*>(&hs[cur][s][k]);
                acc[s] = fmaf(wr[k], hv.x, acc[s]);
                acc[s] = fmaf(wr[k + 1], hv.y, acc[s]);
                acc[s] = fmaf(wr[k + 2], hv.z, acc[s]);
                acc[s] = fmaf(wr[k + 3], hv.w, acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < SEQ; ++s) {
            const float act = (gate == 2) ? tanhf(acc[s]) : sigmoidf_(acc[s]);
            const float fg = __shfl_sync(0xffffffffu, act, e + 8);
            const float gg = __shfl_sync(0xffffffffu, act, e + 16);
            const float og = __shfl_sync(0xffffffffu, act, e + 24);
            if (gate == 0) {
                const float cnew = fg * c[s] + act * gg;
                c[s] = cnew;
                const float hnew = og * tanhf(cnew);
                hs[cur ^ 1][s][unit] = hnew;
                if ((s0 + s) < S) y[((long long)(s0 + s) * T + tt) * (2 * H) + dir * H + unit] = hnew;
            }
        }
        __syncthreads();
    }
    if (gate == 0) {
#pragma unroll
        for (int s = 0; s < SEQ; ++s) {
            if ((s0 + s) >= S) continue;
            if (hn) hn[((long long)dir * S + s0 + s) * H + unit] = hs[T & 1][s][unit];
            if (cn) cn[((long long)dir * S + s0 + s) * H + unit] = c[s];
        }
    }
}

}  // namespace

void launch_lstm_small(const float* gx, const float* whh, const float* h0, const float* c0, float* y, float* hn,
                       float* cn, int S, int T, cudaStream_t st) {
    if (S <= 0 || T <= 0) return;
    dim3 grid((S + SEQ - 1) / SEQ, 2);
    MMEGO_LAUNCH(lstm_small_kernel, grid, dim3(NTH), 0, st, gx, whh, h0, c0, y, hn, cn, S, T);
}

}  // namespace mmego
