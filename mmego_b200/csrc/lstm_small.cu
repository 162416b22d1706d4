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
#include "mma_frag.cuh"
#include "point_layout.h"

namespace mmego {

namespace {

constexpr int H = kSmallH;     // 64
constexpr int NTH = 256;       // 4 gates x 64 units
constexpr int SEQ = 8;         // sequences per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// Branch-free variant for the mma.sync kernel.  The IEEE division above carries a slow path (a CALL taken per lane when
// the denominator leaves [2^-126, 2^126], i.e. for pre-activations beyond +-87, which rnn_pk's first layer reaches);
// taken divergently right before the next warp-wide mma.sync it left the warp unconverged on the B200 and every later
// result of that warp was garbage.  __fdividef is MUFU.RCP + FMUL, 2 ulp, and returns 0 for denominators >= 2^126.
__device__ __forceinline__ float sigmoid_nb(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }

#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
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


#endif  // MMEGO_FFMA_GEN

// ================================================================================================================
// Tensor-core version (default): both GEMMs of the layer on mma.sync m16n8k16 fragments (mma_frag.cuh, fp16 hi/lo
// split products, fp32 accumulation).  Gate columns are interleaved per 8 units (n-tile 4u + gate), see
// pack_small_lstm_mma.
//   lstm_proj_mma_kernel<KS>: gx[M, 512] = x[M, In] W_ih^T + b for all timesteps and both directions.  A CTA owns a
//     slab of 128 gate columns (its fragments staged once in shared memory) and walks over 64-row tiles; a warp holds
//     the A fragments of its 16 rows in registers for the whole slab.
//   lstm_rec_mma_kernel: the recurrence.  A warp owns 16 sequences of one direction for all T steps: h_{t-1} lives in
//     registers AS the A fragments of the next step (the C fragment of unit group u is half of k-step u/2), c in
//     registers, W_hh fragments in shared memory (64 KB).  No block-level synchronisation inside the time loop.
// ================================================================================================================
constexpr int PROJ_NT = 128;
template <int KS>
__global__ void __launch_bounds__(PROJ_NT) lstm_proj_mma_kernel(const float* __restrict__ x, long long ldx,
                                                                const float* __restrict__ blob,
                                                                float* __restrict__ gx, long long M) {
    MMEGO_DYN_SMEM(uint4, wf);                       // [KS][16 n-tiles][32 lanes]
    __shared__ float sbias[128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int slab = blockIdx.y, dir = slab >> 1, jbase = (slab & 1) * 16;
    const size_t ihw = (size_t)mma_frag_words(KS, 32), hhw = (size_t)mma_frag_words(4, 32);
    const uint4* src = reinterpret_cast<const uint4*>(blob + dir * ihw);
    for (int i = tid; i < KS * 16 * 32; i += PROJ_NT) {
        const int s = i / 512, r = i % 512;
        wf[i] = src[(s * 32 + jbase) * 32 + r];
    }
    sbias[tid] = blob[2 * ihw + dir * 256 + jbase * 8 + tid];
    const float os = blob[2 * ihw + 512 + 2 * hhw + dir];
    __syncthreads();
    for (long long r0 = ((long long)blockIdx.x * 4 + warp) * 16; r0 < M; r0 += (long long)gridDim.x * 64) {
        uint32_t ah[KS][4], al[KS][4];
        const bool live0 = r0 + g < M, live1 = r0 + g + 8 < M;
        const float* x0 = x + (r0 + g) * ldx;
        const float* x1 = x + (r0 + g + 8) * ldx;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int col = 16 * s + 8 * hh + 2 * tq;
                const float2 v0 = live0 ? *reinterpret_cast<const float2*>(x0 + col) : make_float2(0.f, 0.f);
                const float2 v1 = live1 ? *reinterpret_cast<const float2*>(x1 + col) : make_float2(0.f, 0.f);
                frag::split2(v0.x, v0.y, ah[s][2 * hh], al[s][2 * hh]);
                frag::split2(v1.x, v1.y, ah[s][2 * hh + 1], al[s][2 * hh + 1]);
            }
        }
        float* o0 = gx + (r0 + g) * 512 + slab * 128 + 2 * tq;
        float* o1 = gx + (r0 + g + 8) * 512 + slab * 128 + 2 * tq;
#pragma unroll 1
        for (int jg = 0; jg < 4; ++jg) {
            float out[4][4];
            frag::dense_tile<KS, 4, false, 16>(wf, sbias, os, ah, al, out, lane, 4 * jg);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (live0) *reinterpret_cast<float2*>(o0 + 8 * (4 * jg + j)) = make_float2(out[j][0], out[j][1]);
                if (live1) *reinterpret_cast<float2*>(o1 + 8 * (4 * jg + j)) = make_float2(out[j][2], out[j][3]);
            }
        }
    }
}

__device__ __forceinline__ float2 ld2_or_zero(const float* p, bool live) {
    return live ? *reinterpret_cast<const float2*>(p) : make_float2(0.f, 0.f);
}

// gx [S][T][2][256] (mma column order), blob as packed by pack_small_lstm_mma (KS = In/16), h0/c0/hn/cn [2][S][64].
// TPC tiles of 16 sequences per CTA; four warps per tile, warp qw owning unit groups 2qw, 2qw+1 (16 hidden units = one
// k-step of the next step's A operand).  h_t travels between the four warps of a tile as ready-made A-fragment words
// through a double-buffered 4 KB shared-memory slot: one __syncthreads per timestep.
template <int TPC>
__global__ void __launch_bounds__(TPC * 128) lstm_rec_mma_kernel(const float* __restrict__ gx,
                                                                 const float* __restrict__ blob, int KS,
                                                                 const float* __restrict__ h0,
                                                                 const float* __restrict__ c0, float* __restrict__ y,
                                                                 float* __restrict__ hn, float* __restrict__ cn, int S,
                                                                 int T) {
    MMEGO_DYN_SMEM(uint4, wf);                       // [4][32 n-tiles][32 lanes] | hx [2][TPC][4 k-steps][hi, lo][32 lanes]
    uint4* hx = wf + 4 * 32 * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int tile = warp >> 2, qw = warp & 3;
    const int dir = blockIdx.y;
    const size_t ihw = (size_t)mma_frag_words(KS, 32), hhw = (size_t)mma_frag_words(4, 32);
    {
        const uint4* src = reinterpret_cast<const uint4*>(blob + 2 * ihw + 512 + dir * hhw);
        for (int i = tid; i < 4 * 32 * 32; i += TPC * 128) wf[i] = src[i];
    }
    const float os = blob[2 * ihw + 512 + 2 * hhw + 2 + dir];
    auto slot = [&](int buf, int s, int plane) { return hx + (((buf * TPC + tile) * 4 + s) * 2 + plane) * 32 + lane; };
    const long long q0 = ((long long)blockIdx.x * TPC + tile) * 16 + g, q1 = q0 + 8;   // this lane's two sequences
    const bool live0 = q0 < S, live1 = q1 < S;
    float c[2][4], hl[2][4];
    {
        uint32_t nh[4], nl[4];
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
            const int col = 8 * (2 * qw + uu) + 2 * tq;
            const float2 ca = ld2_or_zero(c0 ? c0 + ((long long)dir * S + q0) * H + col : nullptr, live0 && c0);
            const float2 cb = ld2_or_zero(c0 ? c0 + ((long long)dir * S + q1) * H + col : nullptr, live1 && c0);
            c[uu][0] = ca.x; c[uu][1] = ca.y; c[uu][2] = cb.x; c[uu][3] = cb.y;
            const float2 ha = ld2_or_zero(h0 ? h0 + ((long long)dir * S + q0) * H + col : nullptr, live0 && h0);
            const float2 hb = ld2_or_zero(h0 ? h0 + ((long long)dir * S + q1) * H + col : nullptr, live1 && h0);
            hl[uu][0] = ha.x; hl[uu][1] = ha.y; hl[uu][2] = hb.x; hl[uu][3] = hb.y;
            frag::split2(ha.x, ha.y, nh[2 * uu], nl[2 * uu]);
            frag::split2(hb.x, hb.y, nh[2 * uu + 1], nl[2 * uu + 1]);
        }
        *slot(0, qw, 0) = make_uint4(nh[0], nh[1], nh[2], nh[3]);
        *slot(0, qw, 1) = make_uint4(nl[0], nl[1], nl[2], nl[3]);
    }
    // gx of the first step (rows g / g+8; [unit group][gate])
    float2 pa[2][4], pb[2][4];
    auto load_gx = [&](int step, float2 (&a)[2][4], float2 (&b)[2][4]) {
        const int tt = dir ? (T - 1 - step) : step;
        const float* g0 = gx + ((q0 * T + tt) * 2 + dir) * 256 + 2 * tq;
        const float* g1 = gx + ((q1 * T + tt) * 2 + dir) * 256 + 2 * tq;
#pragma unroll
        for (int uu = 0; uu < 2; ++uu)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a[uu][k] = ld2_or_zero(g0 + 8 * (4 * (2 * qw + uu) + k), live0);
                b[uu][k] = ld2_or_zero(g1 + 8 * (4 * (2 * qw + uu) + k), live1);
            }
    };
    if (T > 0) load_gx(0, pa, pb);
    __syncthreads();

    for (int step = 0; step < T; ++step) {
        const int buf = step & 1;
        const int tt = dir ? (T - 1 - step) : step;
        float2 na[2][4], nb[2][4];
        if (step + 1 < T) load_gx(step + 1, na, nb);
        uint32_t ah[4][4], al[4][4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint4 vh = *slot(buf, s, 0), vl = *slot(buf, s, 1);
            ah[s][0] = vh.x; ah[s][1] = vh.y; ah[s][2] = vh.z; ah[s][3] = vh.w;
            al[s][0] = vl.x; al[s][1] = vl.y; al[s][2] = vl.z; al[s][3] = vl.w;
        }
        uint32_t nh[4], nl[4];
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
            const int u = 2 * qw + uu;
            float pre[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float big[4] = {0.f, 0.f, 0.f, 0.f}, small[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int s = 0; s < 4; ++s) frag::mma3(big, small, ah[s], al[s], wf[(s * 32 + 4 * u + k) * 32 + lane]);
                pre[k][0] = fmaf(big[0] + small[0], os, pa[uu][k].x);
                pre[k][1] = fmaf(big[1] + small[1], os, pa[uu][k].y);
                pre[k][2] = fmaf(big[2] + small[2], os, pb[uu][k].x);
                pre[k][3] = fmaf(big[3] + small[3], os, pb[uu][k].y);
            }
            float hv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float ig = sigmoid_nb(pre[0][i]), fg = sigmoid_nb(pre[1][i]), gg = tanhf(pre[2][i]),
                            og = sigmoid_nb(pre[3][i]);
                const float cnew = fg * c[uu][i] + ig * gg;
                c[uu][i] = cnew;
                hv[i] = og * tanhf(cnew);
                hl[uu][i] = hv[i];
            }
            float* y0 = y + (q0 * T + tt) * (2 * H) + dir * H + 8 * u + 2 * tq;
            float* y1 = y + (q1 * T + tt) * (2 * H) + dir * H + 8 * u + 2 * tq;
            if (live0) *reinterpret_cast<float2*>(y0) = make_float2(hv[0], hv[1]);
            if (live1) *reinterpret_cast<float2*>(y1) = make_float2(hv[2], hv[3]);
            frag::split2(hv[0], hv[1], nh[2 * uu], nl[2 * uu]);
            frag::split2(hv[2], hv[3], nh[2 * uu + 1], nl[2 * uu + 1]);
        }
        *slot(buf ^ 1, qw, 0) = make_uint4(nh[0], nh[1], nh[2], nh[3]);
        *slot(buf ^ 1, qw, 1) = make_uint4(nl[0], nl[1], nl[2], nl[3]);
#pragma unroll
        for (int uu = 0; uu < 2; ++uu)
#pragma unroll
            for (int k = 0; k < 4; ++k) { pa[uu][k] = na[uu][k]; pb[uu][k] = nb[uu][k]; }
        __syncthreads();
    }
#pragma unroll
    for (int uu = 0; uu < 2; ++uu) {
        const int col = 8 * (2 * qw + uu) + 2 * tq;
        if (hn) {
            if (live0) *reinterpret_cast<float2*>(hn + ((long long)dir * S + q0) * H + col) = make_float2(hl[uu][0], hl[uu][1]);
            if (live1) *reinterpret_cast<float2*>(hn + ((long long)dir * S + q1) * H + col) = make_float2(hl[uu][2], hl[uu][3]);
        }
        if (cn) {
            if (live0) *reinterpret_cast<float2*>(cn + ((long long)dir * S + q0) * H + col) = make_float2(c[uu][0], c[uu][1]);
            if (live1) *reinterpret_cast<float2*>(cn + ((long long)dir * S + q1) * H + col) = make_float2(c[uu][2], c[uu][3]);
        }
    }
}

}  // namespace

#ifdef MMEGO_FFMA_GEN   // fp32 FFMA generation: emulator suite and -DMMEGO_WITH_FFMA test builds only (not in the product library)
void launch_lstm_small(const float* gx, const float* whh, const float* h0, const float* c0, float* y, float* hn,
                       float* cn, int S, int T, cudaStream_t st) {
    if (S <= 0 || T <= 0) return;
    dim3 grid((S + SEQ - 1) / SEQ, 2);
    MMEGO_LAUNCH(lstm_small_kernel, grid, dim3(NTH), 0, st, gx, whh, h0, c0, y, hn, cn, S, T);
}

#endif  // MMEGO_FFMA_GEN

// One H=64 bidirectional layer on mma.sync: x [S*T, In] (row stride ldx) -> gx (workspace [S*T, 512]) -> y [S, T, 128]
void launch_lstm_small_mma(const float* x, long long ldx, int In, const float* blob, float* gx, const float* h0,
                           const float* c0, float* y, float* hn, float* cn, int S, int T, int sm_count,
                           cudaStream_t st) {
    if (S <= 0 || T <= 0) return;
    const int KS = In / 16;
    const long long M = (long long)S * T;
    const size_t psmem = (size_t)KS * 16 * 32 * sizeof(uint4), rsmem = (size_t)4 * 32 * 32 * sizeof(uint4);
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set)) {
        cudaFuncSetAttribute(lstm_proj_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16 * 32 * 16);
        cudaFuncSetAttribute(lstm_proj_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16 * 32 * 16);
        cudaFuncSetAttribute(lstm_proj_mma_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16 * 32 * 16);
        cudaFuncSetAttribute(lstm_rec_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem + 2 * 1 * 8192);
        cudaFuncSetAttribute(lstm_rec_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem + 2 * 2 * 8192);
        cudaFuncSetAttribute(lstm_rec_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem + 2 * 4 * 8192);
    }
    const long long tiles = (M + 63) / 64;
    const long long per_slab = sm_count >= 2 ? sm_count / 2 : 1;   // 4 slabs x sm_count/2 persistent CTAs (twice as many measured 10 % slower)
    dim3 pgrid((unsigned)(tiles < per_slab ? tiles : per_slab), 4);
    if (KS == 4) {
        MMEGO_LAUNCH(lstm_proj_mma_kernel<4>, pgrid, dim3(PROJ_NT), psmem, st, x, ldx, blob, gx, M);
    } else if (KS == 8) {
        MMEGO_LAUNCH(lstm_proj_mma_kernel<8>, pgrid, dim3(PROJ_NT), psmem, st, x, ldx, blob, gx, M);
    } else {
        MMEGO_LAUNCH(lstm_proj_mma_kernel<12>, pgrid, dim3(PROJ_NT), psmem, st, x, ldx, blob, gx, M);
    }
    // tiles (16 sequences) per CTA: as many as still leave about one CTA per SM
    const long long tiles16 = ((long long)S + 15) / 16;
    const int tpc = tiles16 * 2 >= 4LL * sm_count ? 4 : (tiles16 * 2 >= 2LL * sm_count ? 2 : 1);
    const size_t rsm = rsmem + (size_t)2 * tpc * 4 * 2 * 32 * sizeof(uint4);
    dim3 rgrid((unsigned)((tiles16 + tpc - 1) / tpc), 2);
    if (tpc == 4) {
        MMEGO_LAUNCH(lstm_rec_mma_kernel<4>, rgrid, dim3(512), rsm, st, gx, blob, KS, h0, c0, y, hn, cn, S, T);
    } else if (tpc == 2) {
        MMEGO_LAUNCH(lstm_rec_mma_kernel<2>, rgrid, dim3(256), rsm, st, gx, blob, KS, h0, c0, y, hn, cn, S, T);
    } else {
        MMEGO_LAUNCH(lstm_rec_mma_kernel<1>, rgrid, dim3(128), rsm, st, gx, blob, KS, h0, c0, y, hn, cn, S, T);
    }
}

}  // namespace mmego
