// gemm_tc.cu -- the ST-GCN key encoder of Lower_Net (Net/GCN.py:332-355, Net/Lower_Net.py:149-167) on the tensor cores.
//
// Activations are channel-last rows (snippet b, row r = frame*15 + joint) stored as fp16 hi/lo planes [B][RP][C]
// (RP = L*15 rows per snippet) of 2^4 * value.  Every convolution of the network is a row GEMM:
//   graph conv   U  = relu( [Y A_0 | Y A_1] Wg^T + bias[joint] )            (aggregation = gcn_agg_split_kernel)
//   temporal     Y' = relu( sum_tau U[row + (tau-4)*15] Wt_tau^T + Y Wr^T + b )   (9 row-shifted K segments + residual)
//   fcn          E  = Y Wf^T + b, stored as the reference's raw [64][L*15] block (Net/GCN.py:352-353)
// One persistent tcgen05 kernel serves all of them: CTA tile = 128 rows of ONE snippet x all N <= 128 output channels,
// operands by TMA from a 3-D tensor map (channel, row-in-snippet, snippet): a shifted row window that leaves the snippet
// is zero-filled by the TMA unit, which is exactly the (9,1) convolution's temporal zero padding; channels beyond C
// (C = 8 or 32 < the 64-wide K block) are zero-filled the same way.  Precision scheme and warp roles are those of
// lstm_tc.cu (fp16x3 split products, K chunks accumulated in TMEM and drained into fp32 registers).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>

#include "internal.h"
#include "pack.h"
#include "tc_common.cuh"

namespace mmego {

namespace {

using namespace tc;

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + kEpiWarps * 32;   // 640
constexpr int A_TILE = BM * BK * 2;              // 16 KB per plane
constexpr float kGcnActScale = 16.0f, kGcnActInv = 1.0f / 16.0f;
// Every producer of an fp16 hi/lo plane saturates (|2^4 v| <= 65000, i.e. |v| <= ~4062) instead of overflowing to inf
// in the hi plane and NaN in the lo plane (v - inf); the representable range is stated in include/mmego_b200.h.
__device__ __forceinline__ float sat16(float v) { return fminf(fmaxf(v, -65000.f), 65000.f); }

template <int BN>
struct Cfg {
    static constexpr int W_TILE = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * (A_TILE + W_TILE);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
    static constexpr int BIAS_LD = BN + 1;                               // padded: rows r % rowmod hit distinct banks
    static constexpr int BIAS_BYTES = 16 * BIAS_LD * 4;                  // bias table [rowmod <= 16][BN] in shared memory
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + BIAS_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;    // 64, 128, 256: powers of two
    static constexpr int COLS_PER_THREAD = BN / 4;                       // 8, 16, 32
};

struct GemmTcParams {
    int B, RP, rtiles;       // snippets, rows per snippet, row tiles per snippet
    int n0, kb0, kb1;        // # of A0 segments (1 or 9), K blocks per A0 segment, K blocks of the A1 segment (0 = none)
    int shift_step;          // rows between consecutive A0 segments (15 = one frame); segment s is shifted (s - n0/2)*step
    int kb_chunk;
    const float* bias;       // [N] or [rowmod][N]
    int rowmod;
    int relu;
    int N;                   // == BN
    float out_scale;         // 1 / (weight scale * activation scale)
    __half* out_hi;          // planes [B][RP][N] (epilogue 0)
    __half* out_lo;
    float* out_f6;           // fp32 [B][N][RP] (epilogue 1), used when out_hi == nullptr
    int stages, stage_bytes; // shared-memory ring: stages of stage_bytes (A tile pair [+ W tile pair])
    int w_res;               // 1: the whole weight matrix (kb_total K blocks, both planes) is loaded ONCE per CTA into a
                             // resident region behind the ring and the stages carry activations only -- the graph convs
                             // and fcn were bound by L2 -> shared-memory traffic, a third of it the same weight tiles
                             // re-loaded for every 128-row tile
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mA0hi, const __grid_constant__ CUtensorMap mA0lo,
               const __grid_constant__ CUtensorMap mA1hi, const __grid_constant__ CUtensorMap mA1lo,
               const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo, const GemmTcParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kMaxStages = 6;
    const int ring_bytes = p.stages * p.stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kMaxStages;
    uint64_t* tfull = bars + 2 * kMaxStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;                       // resident weights have landed
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wbar + 1);
    uint8_t* wres = smem + ((ring_bytes + 256 + C::BIAS_BYTES + 1023) & ~1023);   // 1024-aligned: SWIZZLE_128B tiles

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.B * p.rtiles;
    const int kb_a0 = p.n0 * p.kb0;
    const int kb_total = kb_a0 + p.kb1;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mA0hi);
        prefetch_tensormap(&mA0lo);
        prefetch_tensormap(&mWhi);
        prefetch_tensormap(&mWlo);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], kEpiWarps);
        }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, C::TMEM_COLS);
        tmem_relinquish();
    }
    // bias table in shared memory: with rowmod > 0 (the graph conv's per-joint bias) every lane of an epilogue warp needs a
    // different table row, and reading it from global was a 15-sector gather per load that throttled the LSU
    // (ncu: lg_throttle 7.5, 2.6 M requests x 14.7 sectors on the N=128 graph conv)
    float* sbias = reinterpret_cast<float*>(smem + ring_bytes + 256);
    {
        const int rows = p.rowmod > 0 ? p.rowmod : 1;
        for (int i = threadIdx.x; i < rows * BN; i += kThreads) sbias[(i / BN) * C::BIAS_LD + (i % BN)] = p.bias[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
      if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if (p.w_res && blockIdx.x < total_tiles) {
                mbar_expect_tx(wbar, kb_total * 2 * C::W_TILE);
                for (int kbi = 0; kbi < kb_total; ++kbi) {
                    tma_load_2d(wres + kbi * 2 * C::W_TILE, &mWhi, wbar, kbi * BK, 0);
                    tma_load_2d(wres + kbi * 2 * C::W_TILE + C::W_TILE, &mWlo, wbar, kbi * BK, 0);
                }
            }
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int b = tile / p.rtiles, r0 = (tile % p.rtiles) * BM;
                for (int kbi = 0; kbi < kb_total; ++kbi) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], p.w_res ? 2 * A_TILE : C::STAGE_BYTES);
                    uint8_t* sa = smem + stage * p.stage_bytes;
                    uint8_t* sw = sa + 2 * A_TILE;
                    if (kbi < kb_a0) {
                        const int seg = kbi / p.kb0, kb = kbi - seg * p.kb0;
                        const int shift = (seg - p.n0 / 2) * p.shift_step;
                        tma_load_3d(sa, &mA0hi, &full[stage], kb * BK, r0 + shift, b);
                        tma_load_3d(sa + A_TILE, &mA0lo, &full[stage], kb * BK, r0 + shift, b);
                    } else {
                        const int kb = kbi - kb_a0;
                        tma_load_3d(sa, &mA1hi, &full[stage], kb * BK, r0, b);
                        tma_load_3d(sa + A_TILE, &mA1lo, &full[stage], kb * BK, r0, b);
                    }
                    if (!p.w_res) {
                        tma_load_2d(sw, &mWhi, &full[stage], kbi * BK, 0);
                        tma_load_2d(sw + C::W_TILE, &mWlo, &full[stage], kbi * BK, 0);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BM, BN, 0 /*fp16*/);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t cc = 0;
            if (p.w_res && blockIdx.x < total_tiles) {
                mbar_wait(wbar, 0);
                tc_fence_after();
            }
            const uint32_t wres_u32 = smem_u32(wres);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int c0 = 0; c0 < kb_total; c0 += p.kb_chunk, ++cc) {
                    const uint32_t buf = cc & 1, bph = (cc >> 1) & 1;
                    mbar_wait(&tempty[buf], bph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * BN;
                    const int c1 = min(kb_total, c0 + p.kb_chunk);
                    for (int kb = c0; kb < c1; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(smem + stage * p.stage_bytes);
                        const uint32_t a_lo = a_hi + A_TILE;
                        const uint32_t w_hi = p.w_res ? wres_u32 + kb * 2 * C::W_TILE : a_hi + 2 * A_TILE;
                        const uint32_t w_lo = w_hi + C::W_TILE;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t da_hi = make_sw128_kmajor_desc(a_hi + k * 32);
                            const uint64_t dw_hi = make_sw128_kmajor_desc(w_hi + k * 32);
                            const uint64_t da_lo = make_sw128_kmajor_desc(a_lo + k * 32);
                            const uint64_t dw_lo = make_sw128_kmajor_desc(w_lo + k * 32);
                            mma_f16_ss(d_tmem, da_hi, dw_lo, idesc, (kb > c0) || (k > 0));
                            mma_f16_ss(d_tmem, da_lo, dw_hi, idesc, 1);
                            mma_f16_ss(d_tmem, da_hi, dw_hi, idesc, 1);
                        }
                        mma_commit(&empty[stage]);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    mma_commit(&tfull[buf]);
                }
            }
        }
        __syncwarp();
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // ===================================================================== epilogue (16 warps)
        constexpr int CPT = C::COLS_PER_THREAD;
        const int q = warp & 3;
        const int part = (warp - 4) >> 2;
        const int col0 = part * CPT;
        uint32_t cc = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int b = tile / p.rtiles, r0 = (tile % p.rtiles) * BM;
            float acc[CPT];
            for (int c0 = 0; c0 < kb_total; c0 += p.kb_chunk, ++cc) {
                const uint32_t buf = cc & 1, bph = (cc >> 1) & 1;
                mbar_wait(&tfull[buf], bph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col0);
#pragma unroll
                for (int g = 0; g < CPT / 8; ++g) {
                    uint32_t r[8];
                    tmem_ld_x8(taddr + g * 8, r);
                    tmem_ld_wait();
                    if (c0 == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[g * 8 + j] = __uint_as_float(r[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[g * 8 + j] += __uint_as_float(r[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
            }
            const int r = r0 + q * 32 + lane;          // row inside the snippet
            if (r < p.RP) {
                const float* bias = sbias + (p.rowmod > 0 ? (r % p.rowmod) * C::BIAS_LD : 0) + col0;
                float v[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    float x = fmaf(acc[j], p.out_scale, bias[j]);
                    v[j] = p.relu ? fmaxf(x, 0.f) : x;
                }
                if (p.out_hi) {
                    const long long o = ((long long)b * p.RP + r) * p.N + col0;
#pragma unroll
                    for (int g = 0; g < CPT / 8; ++g) {
                        uint32_t ph[4], pl[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float v0 = sat16(v[g * 8 + 2 * j] * kGcnActScale);
                            const float v1 = sat16(v[g * 8 + 2 * j + 1] * kGcnActScale);
                            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                            const __half l0 = __float2half_rn(v0 - __half2float(h0));
                            const __half l1 = __float2half_rn(v1 - __half2float(h1));
                            ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                            pl[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                        }
                        *reinterpret_cast<uint4*>(p.out_hi + o + g * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                        *reinterpret_cast<uint4*>(p.out_lo + o + g * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    }
                } else {
                    float* dst = p.out_f6 + ((long long)b * p.N + col0) * p.RP + r;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) dst[(long long)j * p.RP] = v[j];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ================================================================================================================
// Snippet-resident temporal convolution (the 64 % of the ST-GCN chain's time):
//     Y'[c', (l, w)] = relu( sum_tau sum_c Wt_tau[c', c] U[c, (l + tau - 4, w)] + sum_c Wr[c', c] Y[c, (l, w)] + b[c'] )
// TRANSPOSED with respect to gemm_tc_kernel above: the output CHANNELS are the MMA's M (the weight tile is the A operand,
// zero-filled by TMA to 128 rows), the snippet's ROWS are its N.  One CTA owns a whole snippet:
//   * per 64-channel block the snippet's activation window (L frames x 16 rows x 64 ch, both planes: 80 KB at L = 20) is
//     loaded ONCE -- a 4-D TMA box over [snippet][frame][joint][channel] whose 16th joint is out of bounds, so the TMA
//     unit pads every frame to 16 rows: a temporal shift of one frame is then 2048 bytes = two swizzle atoms, and the nine
//     taps read the SAME window through UMMA descriptors with a different start address (B operand) and a different
//     column range of the accumulator; rows that would come from outside the snippet are simply not part of the MMA's N
//     range, which IS the convolution's zero padding (no MMA work is spent on it).  The 16th row of a frame only ever
//     feeds the 16th column of a frame, which is never stored.
//   * a weight tile (128 x 64, both planes: 32 KB) is streamed once per (tap, channel block) and feeds N = 16 L columns
//     (320 at L = 20) instead of a 128-row tile: per snippet 0.8 MB cross L2 -> SM for the widest layer instead of 3.7 MB
//     (three 128-row tiles, each re-loading nine shifted windows and the full weight matrix).
//   * no rows are wasted: the row-tiled kernel computes 384 rows for a 300-row snippet.
// Accumulator: one TMEM buffer of 16 L fp32 columns; after each activation window (9 taps x 12 MMAs per k-step) the 16
// epilogue warps drain it into fp32 registers (round-to-nearest adds, see lstm_tc.cu), 4 L columns per thread.
// ================================================================================================================
constexpr int SN_FR = 16;                        // rows per frame in the window (15 joints + 1 pad row)
constexpr int SN_LMAX = 20;                      // frames per snippet supported (accumulators per epilogue thread: 4 L <= 80)
constexpr int SN_WIN_PLANE = SN_LMAX * SN_FR * BK * 2;     // 40 KB
constexpr int SN_WIN = 2 * SN_WIN_PLANE;                   // hi + lo
constexpr int SN_WT_PLANE = 128 * BK * 2;                  // 16 KB
constexpr int SN_WT = 2 * SN_WT_PLANE;
constexpr int SN_NWIN = 2;
constexpr int SN_WRING = 2 * SN_WT;                        // 64 KB of weight tiles: 2 slots of 128 rows, 4 of 64, 8 of 32
constexpr int SN_MAXWT = 8;
constexpr int SN_SMEM = SN_WRING + SN_NWIN * SN_WIN + 1024 + 256;

struct TconvSnipParams {
    int B, L;
    int kbu;                 // 64-channel blocks of U (1 or 2)
    int Cout;                // real output channels (32 / 64 / 128)
    int nwt;                 // weight-ring slots: a slot holds the Cout real rows of a tile, both planes (the MMA reads 128 rows:
                             // the rows beyond belong to the next slot / the window area and only feed output lanes >= Cout)
    int subdrain;            // drain the accumulator twice per U block (taps with shift <= 0, then shift > 0): see the kernel
    int wk;                  // channels per weight tile: 64 (rows of 128 bytes, SWIZZLE_128B) or 32 (rows of 64 bytes, SWIZZLE_64B)
    int chunk256;            // split a tap's N range as (256, rest) instead of two near-equal halves
    int ksu, ksy;            // 16-channel k-steps that hold real channels in a U block (2 or 4) / in the Y block (1..4): the rest
                             // of a 64-channel block is TMA zero fill and is not multiplied
    const float* bias;       // [Cout]
    float out_scale;
    __half* out_hi;          // planes [B][L*15][Cout]
    __half* out_lo;
    unsigned* error;         // set when a wait gave up (the kernel then finishes without hanging)
};

// taps are processed with the zero shift first: it covers every column of the window's accumulator and clears it
__device__ __forceinline__ int snip_tap(int tj) { return tj == 0 ? 4 : (tj <= 4 ? tj - 1 : tj); }

// bounded mbarrier wait: a protocol bug must not hang the GPU box.  test_wait never suspends the thread (try_wait may, for
// a system-dependent time), so the bound below is a real one: ~2^21 polls of >= 32 ns, about 0.1 s; a waiter also leaves
// as soon as anybody else has given up.
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, unsigned* error) {
    for (unsigned i = 0; i < (1u << 21); ++i) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
        if (i > 64) __nanosleep(32);
        if ((i & 255u) == 255u && *reinterpret_cast<volatile unsigned*>(error) != 0u) return false;
    }
    *error = 1u;
    return false;
}

__global__ void __launch_bounds__(kThreads, 1)
tconv_snip_kernel(const __grid_constant__ CUtensorMap mUhi, const __grid_constant__ CUtensorMap mUlo,
                  const __grid_constant__ CUtensorMap mYhi, const __grid_constant__ CUtensorMap mYlo,
                  const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo,
                  const TconvSnipParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* swt = smem;                                    // weight ring: [nwt][hi rows | lo rows]
    uint8_t* swin = smem + SN_WRING;                        // [SN_NWIN][hi | lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(swin + SN_NWIN * SN_WIN);
    uint64_t* wfull = bars;                  // window slot filled
    uint64_t* wempty = bars + SN_NWIN;       // window slot consumed
    uint64_t* tfullb = wempty + SN_NWIN;     // weight slot filled
    uint64_t* temptyb = tfullb + SN_MAXWT;   // weight slot consumed
    uint64_t* dfull = temptyb + SN_MAXWT;    // accumulator complete (a drain group's MMAs done)
    uint64_t* dempty = dfull + 1;            // accumulator drained
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(dempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwin = p.kbu + 1;              // windows per snippet: residual Y first, then the U channel blocks
    // Weight tiles are HALF K blocks (32 channels = 2 k-steps; rows of 64 bytes, SWIZZLE_64B): the 64 KB ring then holds
    // 4 tiles of the 128-channel layer instead of 2, i.e. three loads in flight behind the one being multiplied (with
    // two slots the MMA thread waited for L2 on every tile: tensor pipe 40 % active, profiles/r02_gcn_snip_ncu.txt).
    const int wt_plane = p.Cout * p.wk * 2;  // bytes of one plane of a weight tile (real rows only)
    const int wt_slot = 2 * wt_plane;
    const int kpt = p.wk / 16;               // k-steps per weight tile (2 or 4)
    const int nhu = (p.ksu + kpt - 1) / kpt, nhy = (p.ksy + kpt - 1) / kpt;      // tiles per tap (U) / of the residual block (Y)
    // Drain groups (the accumulator is drained into fp32 registers after each: long accumulation chains in TMEM lose
    // accuracy -- one group per channel block measured 9.5e-6 relative error on the reference vector against 1.5e-6 of
    // the row-tiled kernel): per U block the taps with shift <= 0 (the zero shift first: it covers every column and
    // clears the accumulator), then the taps with shift > 0 (the first of them, shift +1, clears frames 0 .. L-2; the
    // last frame's columns keep stale values that the epilogue skips).  The one-tap residual window joins the first group.
    // Measured (B200, reference vector gcn2.npz / 2048 snippets): one group per block 9.5e-6 and 1.60 ms, two groups
    // 5.6e-6 and 1.79 ms, the row-tiled kernel 1.5e-6 and 1.81 ms; joints stay at 1.5e-6 m either way (tolerance 1e-5 m),
    // so the split (`subdrain`, option gcn_snip bit 4) is off by default.
    const int tiles_a = p.subdrain ? 5 * nhu : (1 << 20);   // tiles of a U window's first group (taps 4, 0, 1, 2, 3)

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mUhi);
        prefetch_tensormap(&mUlo);
        prefetch_tensormap(&mYhi);
        prefetch_tensormap(&mYlo);
        prefetch_tensormap(&mWhi);
        prefetch_tensormap(&mWlo);
        for (int s = 0; s < SN_NWIN; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], 1);
        }
        for (int s = 0; s < SN_MAXWT; ++s) {
            mbar_init(&tfullb[s], 1);
            mbar_init(&temptyb[s], 1);
        }
        mbar_init(dfull, 1);
        mbar_init(dempty, kEpiWarps);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
      if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            uint32_t wi = 0, ti = 0;         // running window / weight-tile counters
            bool ok = true;
            for (int b = blockIdx.x; b < p.B && ok; b += gridDim.x) {
                for (int w = 0; w < nwin && ok; ++w, ++wi) {
                    const int ws = wi % SN_NWIN;
                    ok = mbar_wait_bounded(&wempty[ws], ((wi / SN_NWIN) & 1) ^ 1, p.error);
                    if (!ok) break;
                    uint8_t* dst = swin + ws * SN_WIN;
                    mbar_expect_tx(&wfull[ws], 2 * p.L * SN_FR * BK * 2);
                    if (w == 0) {
                        tma_load_4d(dst, &mYhi, &wfull[ws], 0, 0, 0, b);
                        tma_load_4d(dst + SN_WIN_PLANE, &mYlo, &wfull[ws], 0, 0, 0, b);
                    } else {
                        tma_load_4d(dst, &mUhi, &wfull[ws], (w - 1) * BK, 0, 0, b);
                        tma_load_4d(dst + SN_WIN_PLANE, &mUlo, &wfull[ws], (w - 1) * BK, 0, 0, b);
                    }
                    const int ntile = w == 0 ? nhy : 9 * nhu;
                    for (int tl = 0; tl < ntile && ok; ++tl, ++ti) {
                        const int ts = ti % p.nwt;
                        ok = mbar_wait_bounded(&temptyb[ts], ((ti / p.nwt) & 1) ^ 1, p.error);
                        if (!ok) break;
                        // packed K order of the weights: [tap 0: U blocks][tap 1: ...] ... [residual block]
                        const int kcol = (w == 0 ? 9 * p.kbu * BK + tl * p.wk
                                                 : (snip_tap(tl / nhu) * p.kbu + (w - 1)) * BK + (tl % nhu) * p.wk);
                        uint8_t* wd = swt + ts * wt_slot;
                        mbar_expect_tx(&tfullb[ts], wt_slot);
                        tma_load_2d(wd, &mWhi, &tfullb[ts], kcol, 0);
                        tma_load_2d(wd + wt_plane, &mWlo, &tfullb[ts], kcol, 0);
                    }
                }
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            uint32_t wi = 0, ti = 0, di = 0;
            bool ok = true;
#ifdef MMEGO_DEBUG_SWITCHES
            long long c_total = clock64(), c_d = 0, c_w = 0, c_t = 0, c0;
#define SNIP_T0() c0 = clock64()
#define SNIP_T1(acc) acc += clock64() - c0
#else
#define SNIP_T0()
#define SNIP_T1(acc)
#endif
            for (int b = blockIdx.x; b < p.B && ok; b += gridDim.x) {
                for (int w = 0; w < nwin && ok; ++w, ++wi) {
                    const int ws = wi % SN_NWIN;
                    if (w != 1) {            // start of a drain group: the accumulator must have been drained
                        SNIP_T0();
                        ok = mbar_wait_bounded(dempty, (di & 1) ^ 1, p.error);
                        SNIP_T1(c_d);
                        if (!ok) break;
                    }
                    SNIP_T0();
                    ok = mbar_wait_bounded(&wfull[ws], (wi / SN_NWIN) & 1, p.error);
                    SNIP_T1(c_w);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t win_hi = smem_u32(swin + ws * SN_WIN), win_lo = win_hi + SN_WIN_PLANE;
                    const int ntile = w == 0 ? nhy : 9 * nhu;
                    // tap order (snip_tap): the zero-shift tap first -- it covers every column and clears the accumulator
                    for (int tl = 0; tl < ntile && ok; ++tl, ++ti) {
                        if (w > 0 && tl == tiles_a) {          // second group of this U window
                            mma_commit(dfull);
                            ++di;
                            SNIP_T0();
                            ok = mbar_wait_bounded(dempty, (di & 1) ^ 1, p.error);
                            SNIP_T1(c_d);
                            if (!ok) break;
                        }
                        const int ts = ti % p.nwt;
                        SNIP_T0();
                        ok = mbar_wait_bounded(&tfullb[ts], (ti / p.nwt) & 1, p.error);
                        SNIP_T1(c_t);
                        if (!ok) break;
                        tc_fence_after();
                        const uint32_t w_hi = smem_u32(swt + ts * wt_slot), w_lo = w_hi + wt_plane;
                        const int tap = w == 0 ? 4 : snip_tap(tl / nhu);  // same order as the producer; shift = tap - 4
                        const int half = w == 0 ? tl : tl % nhu;          // which part of the 64-channel block this tile covers
                        const int ks_real = w == 0 ? p.ksy : p.ksu;
                        const int nk = ks_real - kpt * half < kpt ? ks_real - kpt * half : kpt;
                        const int sh = tap - 4;
                        const int f0 = sh < 0 ? -sh : 0, f1 = sh > 0 ? p.L - sh : p.L;     // output frames [f0, f1)
                        if (f1 > f0) {
                            // The tap's N range is cut into (at most) two chunks = two INDEPENDENT accumulator column ranges;
                            // their MMAs are issued alternately, so that consecutive MMAs never accumulate into the same
                            // columns (a chain of dependent accumulations is latency-bound: 142 clk per M128 x N160 x K16
                            // MMA against 75 when the pipe is fed independent work).
                            const int total = (f1 - f0) * SN_FR;
                            const int n0 = total <= 256 ? total : (p.chunk256 ? 256 : ((total / 2 + 15) / 16) * 16);
                            const int nch = total > n0 ? 2 : 1;
                            // everything per chunk is prepared once per tap; per MMA the issuing thread only adds the k-step
                            // offset (32 bytes = 2 descriptor units) to two descriptors
                            uint32_t idesc[2], d_tmem[2];
                            uint64_t bhi[2], blo[2];
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int c0 = c * n0, n = c == 0 ? n0 : total - n0;
                                idesc[c] = make_idesc_f16(128, n > 0 ? n : 16, 0 /*fp16*/);
                                d_tmem[c] = tmem_base + (uint32_t)(f0 * SN_FR + c0);
                                const uint32_t boff = (uint32_t)((f0 + sh) * SN_FR + c0) * 128u + (uint32_t)(kpt * half) * 32u;
                                bhi[c] = make_sw128_kmajor_desc(win_hi + boff);
                                blo[c] = make_sw128_kmajor_desc(win_lo + boff);
                            }
                            const uint64_t ahi = p.wk == 32 ? make_sw64_kmajor_desc(w_hi) : make_sw128_kmajor_desc(w_hi);
                            const uint64_t alo = p.wk == 32 ? make_sw64_kmajor_desc(w_lo) : make_sw128_kmajor_desc(w_lo);
                            const bool clear = (w != 1 && tl == 0) || (w > 0 && tl == tiles_a);   // first tile of a drain group (block 0 adds to the residual)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < nk) {
                                    const uint64_t ko = (uint64_t)(2 * k);
                                    if (k == 0 && clear) {
                                        mma_f16_ss(d_tmem[0], ahi, blo[0], idesc[0], 0);
                                        if (nch == 2) mma_f16_ss(d_tmem[1], ahi, blo[1], idesc[1], 0);
                                    } else {
                                        mma_f16_ss_acc(d_tmem[0], ahi + ko, blo[0] + ko, idesc[0]);
                                        if (nch == 2) mma_f16_ss_acc(d_tmem[1], ahi + ko, blo[1] + ko, idesc[1]);
                                    }
                                    mma_f16_ss_acc(d_tmem[0], alo + ko, bhi[0] + ko, idesc[0]);
                                    if (nch == 2) mma_f16_ss_acc(d_tmem[1], alo + ko, bhi[1] + ko, idesc[1]);
                                    mma_f16_ss_acc(d_tmem[0], ahi + ko, bhi[0] + ko, idesc[0]);
                                    if (nch == 2) mma_f16_ss_acc(d_tmem[1], ahi + ko, bhi[1] + ko, idesc[1]);
                                }
                            }
                        }
                        mma_commit(&temptyb[ts]);
                    }
                    mma_commit(&wempty[ws]);
                    if (w >= 1) {            // end of a drain group
                        mma_commit(dfull);
                        ++di;
                    }
                }
            }
#ifdef MMEGO_DEBUG_SWITCHES
            {   // MMA-thread cycle accounting (test builds): [2] total, [3] waiting for drains, [4] for windows, [5] for weights
                unsigned long long* st = reinterpret_cast<unsigned long long*>(p.error);
                atomicAdd(st + 2, (unsigned long long)(clock64() - c_total));
                atomicAdd(st + 3, (unsigned long long)c_d);
                atomicAdd(st + 4, (unsigned long long)c_w);
                atomicAdd(st + 5, (unsigned long long)c_t);
                atomicAdd(st + 6, 1ull);
            }
#endif
        }
        __syncwarp();
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // ===================================================================== epilogue (16 warps)
        constexpr int CPT = 4 * SN_LMAX;          // accumulator columns per thread (80): part x CPT .. + 4 L
        const int q = warp & 3;                   // TMEM lane quarter = channels 32 q .. 32 q + 31
        const int part = (warp - 4) >> 2;         // quarter of the columns
        const int ch = q * 32 + lane;
        const int cpp = 4 * p.L;                  // live columns per part (multiple of 4; L frames x 16 rows / 4 parts)
        const float bias = ch < p.Cout ? p.bias[ch] : 0.f;
        uint32_t di = 0;
        bool ok = true;
        for (int b = blockIdx.x; b < p.B && ok; b += gridDim.x) {
            float acc[CPT];
            const int ngroups = p.subdrain ? 2 * p.kbu : p.kbu;
            for (int w = 0; w < ngroups && ok; ++w, ++di) {        // one pass per drain group; with subdrain odd groups = taps with shift > 0
                ok = mbar_wait_bounded(dfull, di & 1, p.error);
                if (!ok) break;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * cpp);
#pragma unroll
                for (int g8 = 0; g8 < CPT / 8; ++g8) {
                    if (g8 * 8 < cpp) {           // warp-uniform
                        uint32_t r[8];
                        tmem_ld_x8(taddr + g8 * 8, r);
                        tmem_ld_wait();
                        if (w == 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[g8 * 8 + j] = __uint_as_float(r[j]);
                        } else if (p.subdrain && (w & 1)) {       // shift > 0 taps never write the last frame's columns: stale there
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (((part * cpp + g8 * 8 + j) >> 4) < p.L - 1) acc[g8 * 8 + j] += __uint_as_float(r[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[g8 * 8 + j] += __uint_as_float(r[j]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dempty);
            }
            if (!ok) break;
            if (ch < p.Cout) {
                __half* oh = p.out_hi + (long long)b * p.L * kGcnV * p.Cout + ch;
                __half* ol = p.out_lo + (long long)b * p.L * kGcnV * p.Cout + ch;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int col = part * cpp + j;                   // padded row: frame * 16 + joint
                    const int jt = col & (SN_FR - 1);
                    if (j < cpp && jt < kGcnV) {
                        const int row = (col >> 4) * kGcnV + jt;
                        const float v = sat16(fmaxf(fmaf(acc[j], p.out_scale, bias), 0.f) * kGcnActScale);
                        const __half hh = __float2half_rn(v);
                        oh[(long long)row * p.Cout] = hh;
                        ol[(long long)row * p.Cout] = __float2half_rn(v - __half2float(hh));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------- memory-bound glue
// Lower_Net front (Net/Lower_Net.py:229, Net/GCN.py:339-344): Transform2H of the upper-body joints + data_bn.
//   upper [F,15,3] -> uh [F,45] fp32 (also the 45 extra inputs of fusion.fc0), y0 planes [F*15][8] (3 channels + zeros)
__global__ void gcn_prep_split_kernel(const float* __restrict__ upper, const float* __restrict__ R,
                                      const float* __restrict__ t, const float* __restrict__ bn, float* __restrict__ uh,
                                      __half* __restrict__ yhi, __half* __restrict__ ylo, long long F) {
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
    __half hi[8], lo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float val = 0.f;
        if (c < 3) {
            uh[i * 3 + c] = h[c];
            val = sat16((h[c] * bn[v * 3 + c] + bn[45 + v * 3 + c]) * kGcnActScale);
        }
        hi[c] = __float2half_rn(val);
        lo[c] = __float2half_rn(val - __half2float(hi[c]));
    }
    *reinterpret_cast<uint4*>(yhi + i * 8) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(ylo + i * 8) = *reinterpret_cast<const uint4*>(lo);
}

// standalone GCN.Model.extract_feature entry: x [B,3,T,15] -> y0 planes with data_bn applied
__global__ void gcn_prep_raw_split_kernel(const float* __restrict__ x, const float* __restrict__ bn,
                                          __half* __restrict__ yhi, __half* __restrict__ ylo, int B, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, t, v)
    const long long total = (long long)B * T * kGcnV;
    if (i >= total) return;
    const int v = (int)(i % kGcnV);
    const long long bt = i / kGcnV;
    const int tt = (int)(bt % T);
    const long long b = bt / T;
    __half hi[8], lo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float val = 0.f;
        if (c < 3) val = sat16((x[((b * 3 + c) * T + tt) * kGcnV + v] * bn[v * 3 + c] + bn[45 + v * 3 + c]) * kGcnActScale);
        hi[c] = __float2half_rn(val);
        lo[c] = __float2half_rn(val - __half2float(hi[c]));
    }
    *reinterpret_cast<uint4*>(yhi + i * 8) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(ylo + i * 8) = *reinterpret_cast<const uint4*>(lo);
}

// neighbourhood aggregation (einsum of Net/GCN.py:62, commuted in front of the 1x1 conv):
//   ya[(f,w)][k*C + c] = sum_v y[(f,v)][c] * Ahat[k][v][w];   y planes have row stride CS, ya planes row stride OS
__global__ void gcn_agg_split_kernel(const __half* __restrict__ yhi, const __half* __restrict__ ylo,
                                     const float* __restrict__ ahat, __half* __restrict__ ohi, __half* __restrict__ olo,
                                     long long F, int C, int CS, int OS) {
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
        const long long base = f * kGcnV * CS + c;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int v = 0; v < kGcnV; ++v) {
            const float val = __half2float(yhi[base + v * CS]) + __half2float(ylo[base + v * CS]);
            a0 = fmaf(val, sa[v * kGcnV + w], a0);
            a1 = fmaf(val, sa[kGcnV * kGcnV + v * kGcnV + w], a1);
        }
        a0 = sat16(a0);
        a1 = sat16(a1);
        const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);     // values already carry the 2^4 scale
        ohi[fw * OS + c] = h0;
        olo[fw * OS + c] = __float2half_rn(a0 - __half2float(h0));
        ohi[fw * OS + C + c] = h1;
        olo[fw * OS + C + c] = __float2half_rn(a1 - __half2float(h1));
    }
}

// The same aggregation for C % 8 == 0 (layers 1 and 2), written for HBM: the kernel above issues 30 two-byte loads
// per output pair.  Here a CTA stages the 15 x C inputs of a few frames in shared memory (16-byte plane loads, hi + lo
// summed once), then thread (w, 8-channel chunk) runs the dense 15-term contraction for both partitions from shared
// memory and writes its results with four 16-byte stores.  Double-buffered: one barrier per group of frames.
template <int C>
__global__ void __launch_bounds__(256) gcn_agg8_split_kernel(const __half* __restrict__ yhi, const __half* __restrict__ ylo,
                                                             const float* __restrict__ ahat, __half* __restrict__ ohi,
                                                             __half* __restrict__ olo, long long F) {
    constexpr int NCH = C / 8;                  // 8-channel chunks per row
    constexpr int ITEMS = kGcnV * NCH;          // (joint, chunk) pairs per frame
    constexpr int FPB = 256 / ITEMS;            // frames per CTA iteration
    __shared__ float sa[2 * kGcnV * kGcnV];
    __shared__ __align__(16) float ys[2][2][FPB][kGcnV][NCH][4];   // [buffer][first/second half of a chunk]
    for (int i = threadIdx.x; i < 2 * kGcnV * kGcnV; i += 256) sa[i] = ahat[i];
    const int tid = threadIdx.x;
    const bool active = tid < FPB * ITEMS;
    const int fl = tid / ITEMS, it = tid % ITEMS, w = it / NCH, ch = it % NCH;
    const long long ngroups = (F + FPB - 1) / FPB;
    auto stage = [&](long long grp, int buf) {
        const long long f = grp * FPB + fl;
        if (!active) return;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (f < F) {
            const long long o = (f * kGcnV + w) * C + ch * 8;
            const uint4 a = *reinterpret_cast<const uint4*>(yhi + o);
            const uint4 b = *reinterpret_cast<const uint4*>(ylo + o);
            const __half2* ah = reinterpret_cast<const __half2*>(&a);
            const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 x = __half22float2(ah[k]), y = __half22float2(bh[k]);
                v[2 * k] = x.x + y.x;
                v[2 * k + 1] = x.y + y.y;
            }
        }
        *reinterpret_cast<float4*>(ys[buf][0][fl][w][ch]) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(ys[buf][1][fl][w][ch]) = make_float4(v[4], v[5], v[6], v[7]);
    };
    int buf = 0;
    if (blockIdx.x < ngroups) stage(blockIdx.x, 0);
    __syncthreads();
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, buf ^= 1) {
        if (grp + gridDim.x < ngroups) stage(grp + gridDim.x, buf ^ 1);
        const long long f = grp * FPB + fl;
        if (active && f < F) {
            float a0[8], a1[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a0[k] = a1[k] = 0.f;
#pragma unroll
            for (int v = 0; v < kGcnV; ++v) {
                const float c0 = sa[v * kGcnV + w], c1 = sa[kGcnV * kGcnV + v * kGcnV + w];
                const float4 p = *reinterpret_cast<const float4*>(ys[buf][0][fl][v][ch]);
                const float4 q = *reinterpret_cast<const float4*>(ys[buf][1][fl][v][ch]);
                const float x[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a0[k] = fmaf(x[k], c0, a0[k]);
                    a1[k] = fmaf(x[k], c1, a1[k]);
                }
            }
            __align__(16) __half h0[8], l0[8], h1[8], l1[8];     // values already carry the 2^4 scale
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a0[k] = sat16(a0[k]);
                a1[k] = sat16(a1[k]);
                h0[k] = __float2half_rn(a0[k]);
                l0[k] = __float2half_rn(a0[k] - __half2float(h0[k]));
                h1[k] = __float2half_rn(a1[k]);
                l1[k] = __float2half_rn(a1[k] - __half2float(h1[k]));
            }
            const long long o = (f * kGcnV + w) * (2 * C) + ch * 8;
            *reinterpret_cast<uint4*>(ohi + o) = *reinterpret_cast<const uint4*>(h0);
            *reinterpret_cast<uint4*>(olo + o) = *reinterpret_cast<const uint4*>(l0);
            *reinterpret_cast<uint4*>(ohi + o + C) = *reinterpret_cast<const uint4*>(h1);
            *reinterpret_cast<uint4*>(olo + o + C) = *reinterpret_cast<const uint4*>(l1);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        if (q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// planes [B][RP][C] fp16; box = 64 channels x 128 rows x 1 snippet
bool make_rows_map(CUtensorMap* m, const void* base, int B, int RP, int C) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)RP, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)RP * C * 2};
    cuuint32_t box[3] = {BK, BM, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool make_w_map(CUtensorMap* m, const void* base, int rows, int K) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
void launch_gemm_tc(const CUtensorMap& a0h, const CUtensorMap& a0l, const CUtensorMap& a1h, const CUtensorMap& a1l,
                    const CUtensorMap& wh, const CUtensorMap& wl, GemmTcParams p, int w_res_opt, int sm_count, cudaStream_t st) {
    using C = Cfg<BN>;
    constexpr int kBudget = 208 * 1024;                 // ring + resident weights
    const int kb_total = p.n0 * p.kb0 + p.kb1;
    const int wres_bytes = kb_total * 2 * C::W_TILE;
    // resident weights when they leave room for at least three activation-only stages
    p.w_res = (w_res_opt && wres_bytes + 3 * 2 * A_TILE <= kBudget) ? 1 : 0;
    if (p.w_res) {
        p.stage_bytes = 2 * A_TILE;
        p.stages = std::min(6, (kBudget - wres_bytes) / p.stage_bytes);
    } else {
        p.stage_bytes = C::STAGE_BYTES;
        p.stages = C::STAGES;
    }
    const int smem = p.stages * p.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/ + C::BIAS_BYTES + 1024 /*W alignment*/ +
                     (p.w_res ? wres_bytes : 0);
    static int attr_bytes[64] = {0};
    int d = 0;
    cudaGetDevice(&d);
    if (attr_bytes[d & 63] < smem) {
        cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_bytes[d & 63] = smem;
    }
    const int total = p.B * p.rtiles;
    const int grid = total < sm_count ? total : sm_count;
    ++t_launches;
    gemm_tc_kernel<BN><<<grid, kThreads, smem, st>>>(a0h, a0l, a1h, a1l, wh, wl, p);
}

}  // namespace

// ================================================================================================ host interface
// Re-packs a host GEMM (segments padded to 16) into fp16 hi/lo planes with every segment padded to 64 columns.
bool tc_pack_gemm(mmego_handle* h, const HostPackedGemm& g, TcGemmW& out) {
    int K64 = 0;
    for (int s = 0; s < g.nseg; ++s) K64 += (g.k[s] + 63) / 64 * 64;
    std::vector<float> w((size_t)g.N * K64, 0.f);
    float wmax = 0.f;
    for (int n = 0; n < g.N; ++n) {
        int src = 0, dst = 0;
        for (int s = 0; s < g.nseg; ++s) {
            for (int k = 0; k < g.k[s]; ++k) {
                const float v = g.w[(size_t)n * g.ldw + src + k];
                w[(size_t)n * K64 + dst + k] = v;
                wmax = std::max(wmax, std::fabs(v));
            }
            src += g.kpad[s];
            dst += (g.k[s] + 63) / 64 * 64;
        }
    }
    int e = 11;
    while (e > 0 && wmax * std::ldexp(1.0f, e) > 30000.f) --e;
    const float scale = std::ldexp(1.0f, e);
    std::vector<__half> hi(w.size()), lo(w.size());
    for (size_t i = 0; i < w.size(); ++i) {
        const float v = w[i] * scale;
        const __half a = __float2half_rn(v);
        hi[i] = a;
        lo[i] = __float2half_rn(v - __half2float(a));
    }
    cudaSetDevice(h->device);
    auto up = [&](const void* src, size_t bytes, void** dst) {
        if (cudaMalloc(dst, bytes) != cudaSuccess) return false;
        h->owned.push_back(*dst);
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    // a re-pack (load_state_dict, in-place weight edit) replaces the planes: release the previous ones first
    handle_free(h, out.whi);
    handle_free(h, out.wlo);
    handle_free(h, out.bias);
    out.whi = out.wlo = nullptr;
    out.bias = nullptr;
    void *dhi = nullptr, *dlo = nullptr, *db = nullptr;
    if (!up(hi.data(), hi.size() * 2, &dhi) || !up(lo.data(), lo.size() * 2, &dlo) ||
        !up(g.bias.data(), g.bias.size() * 4, &db))
        return false;
    out.whi = dhi;
    out.wlo = dlo;
    out.bias = static_cast<float*>(db);
    out.N = g.N;
    out.K64 = K64;
    out.out_scale = 1.0f / scale;
    return make_w_map(reinterpret_cast<CUtensorMap*>(&out.map_hi), dhi, g.N, K64) &&
           make_w_map(reinterpret_cast<CUtensorMap*>(&out.map_lo), dlo, g.N, K64);
}

// One GEMM of the chain.  a0: planes [B][RP][c0] read through n0 shifted windows; a1 (optional): planes [B][RP][c1].
int tc_gcn_gemm(mmego_handle* h, const TcGemmW& w, const void* a0hi, const void* a0lo, int c0, int n0, const void* a1hi,
                const void* a1lo, int c1, int rowmod, int relu, void* outhi, void* outlo, float* out_f6, int B, int RP,
                cudaStream_t st) {
    CUtensorMap m0h, m0l, m1h, m1l;
    if (!make_rows_map(&m0h, a0hi, B, RP, c0) || !make_rows_map(&m0l, a0lo, B, RP, c0)) return -1;
    if (a1hi) {
        if (!make_rows_map(&m1h, a1hi, B, RP, c1) || !make_rows_map(&m1l, a1lo, B, RP, c1)) return -1;
    } else {
        m1h = m0h;
        m1l = m0l;
    }
    GemmTcParams p{};
    p.B = B;
    p.RP = RP;
    p.rtiles = (RP + BM - 1) / BM;
    p.n0 = n0;
    p.kb0 = (c0 + BK - 1) / BK;
    p.kb1 = a1hi ? (c1 + BK - 1) / BK : 0;
    p.shift_step = kGcnV;
    p.kb_chunk = h->gcn_kb_chunk > 0 ? h->gcn_kb_chunk : (p.n0 * p.kb0 + p.kb1);
    p.bias = w.bias;
    p.rowmod = rowmod;
    p.relu = relu;
    p.N = w.N;
    p.out_scale = w.out_scale * kGcnActInv;
    p.out_hi = static_cast<__half*>(outhi);
    p.out_lo = static_cast<__half*>(outlo);
    p.out_f6 = out_f6;
    if ((p.n0 * p.kb0 + p.kb1) * BK != w.K64) return -2;
    const CUtensorMap& wh = *reinterpret_cast<const CUtensorMap*>(&w.map_hi);
    const CUtensorMap& wl = *reinterpret_cast<const CUtensorMap*>(&w.map_lo);
    const int wr = h->gcn_w_res;
    if (w.N == 128) launch_gemm_tc<128>(m0h, m0l, m1h, m1l, wh, wl, p, wr, h->sm_count, st);
    else if (w.N == 64) launch_gemm_tc<64>(m0h, m0l, m1h, m1l, wh, wl, p, wr, h->sm_count, st);
    else if (w.N == 32) launch_gemm_tc<32>(m0h, m0l, m1h, m1l, wh, wl, p, wr, h->sm_count, st);
    else return -3;
    return 0;
}

// activation planes [B][L][15][C] fp16 as a 4-D tensor; box = 64 channels x 16 joints x L frames x 1 snippet: the 16th
// joint is out of bounds, so every frame arrives padded to 16 rows (zero row), channels beyond C as zeros
static bool make_snip_map(CUtensorMap* m, const void* base, int B, int L, int C) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)kGcnV, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)kGcnV * C * 2, (cuuint64_t)L * kGcnV * C * 2};
    cuuint32_t box[4] = {BK, SN_FR, (cuuint32_t)L, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weights [rows][K] fp16; box = 32 x rows: half a K block, the real rows of a tile, 64-byte rows with SWIZZLE_64B (the
// MMA's M is always 128, see TconvSnipParams::nwt)
static bool make_w128_map(CUtensorMap* m, const void* base, int rows, int K, int wk) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)wk, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, wk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tc_gcn_tconv_snip_supported(int L) { return L >= 1 && L <= SN_LMAX; }

// Temporal conv + residual of one ST-GCN layer, snippet-resident (tconv_snip_kernel): U planes [B][L*15][cu] (9 taps),
// Y planes [B][L*15][cy] (residual 1x1 conv), weights packed as for tc_gcn_gemm -> out planes [B][L*15][w.N].
int tc_gcn_tconv_snip(mmego_handle* h, const TcGemmW& w, const void* uhi, const void* ulo, int cu, const void* yhi,
                      const void* ylo, int cy, int cy_real, void* outhi, void* outlo, int B, int L, int subdrain, cudaStream_t st) {
    if (!tc_gcn_tconv_snip_supported(L) || cu != w.N) return -4;
    const int kbu = (cu + BK - 1) / BK;
    if ((9 * kbu + (cy + BK - 1) / BK) * BK != w.K64 || (cy + BK - 1) / BK != 1) return -2;
    const int wk = (h->gcn_snip & 2) ? 32 : 64;
    CUtensorMap mUh, mUl, mYh, mYl, mWh, mWl;
    if (!make_snip_map(&mUh, uhi, B, L, cu) || !make_snip_map(&mUl, ulo, B, L, cu) || !make_snip_map(&mYh, yhi, B, L, cy) ||
        !make_snip_map(&mYl, ylo, B, L, cy) || !make_w128_map(&mWh, w.whi, w.N, w.K64, wk) || !make_w128_map(&mWl, w.wlo, w.N, w.K64, wk))
        return -1;
    if (!h->dev_error) {
        if (cudaMalloc(reinterpret_cast<void**>(&h->dev_error), 64) != cudaSuccess) return -1;
        cudaMemset(h->dev_error, 0, 64);
    }
    TconvSnipParams p{};
    p.B = B;
    p.L = L;
    p.kbu = kbu;
    p.Cout = w.N;
    p.wk = wk;
    p.chunk256 = (h->gcn_snip & 4) ? 1 : 0;
    p.subdrain = subdrain ? 1 : 0;
    p.nwt = SN_WRING / (2 * w.N * wk * 2);
    if (p.nwt > SN_MAXWT) p.nwt = SN_MAXWT;
    if (w.N % 8 != 0 || w.N > 128 || p.nwt < 2) return -3;
    p.ksu = cu >= BK ? 4 : (cu + 31) / 32 * 2;              // real 16-channel k-steps per U block (rounded to a half block)
    p.ksy = (cy_real + 15) / 16;
    p.bias = w.bias;
    p.out_scale = w.out_scale * kGcnActInv;
    p.out_hi = static_cast<__half*>(outhi);
    p.out_lo = static_cast<__half*>(outlo);
    p.error = h->dev_error;
    const int grid = B < h->sm_count ? B : h->sm_count;
    ++t_launches;
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set))
        cudaFuncSetAttribute(tconv_snip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SN_SMEM);
    tconv_snip_kernel<<<grid, kThreads, SN_SMEM, st>>>(mUh, mUl, mYh, mYl, mWh, mWl, p);
    return 0;
}

void tc_gcn_prep(const float* upper, const float* R, const float* t, const float* bn, float* uh, void* yhi, void* ylo,
                 long long F, cudaStream_t st) {
    const long long total = F * kGcnV;
    if (total <= 0) return;
    ++t_launches;
    gcn_prep_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(upper, R, t, bn, uh, static_cast<__half*>(yhi),
                                                                           static_cast<__half*>(ylo), F);
}
void tc_gcn_prep_raw(const float* x, const float* bn, void* yhi, void* ylo, int B, int T, cudaStream_t st) {
    const long long total = (long long)B * T * kGcnV;
    if (total <= 0) return;
    ++t_launches;
    gcn_prep_raw_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, bn, static_cast<__half*>(yhi),
                                                                               static_cast<__half*>(ylo), B, T);
}
void tc_gcn_agg(const void* yhi, const void* ylo, const float* ahat, void* ohi, void* olo, long long F, int C, int CS,
                int OS, int sm_count, cudaStream_t st) {
    const long long total = F * kGcnV * C;
    if (total <= 0) return;
    const __half* ih = static_cast<const __half*>(yhi);
    const __half* il = static_cast<const __half*>(ylo);
    __half* oh = static_cast<__half*>(ohi);
    __half* ol = static_cast<__half*>(olo);
    ++t_launches;
    if ((C == 32 || C == 64) && CS == C && OS == 2 * C) {
        const int fpb = 256 / (kGcnV * (C / 8));
        long long blocks = (F + fpb - 1) / fpb;
        const long long cap = (long long)sm_count * 8;
        if (blocks > cap) blocks = cap;
        if (C == 32) gcn_agg8_split_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(ih, il, ahat, oh, ol, F);
        else gcn_agg8_split_kernel<64><<<(unsigned)blocks, 256, 0, st>>>(ih, il, ahat, oh, ol, F);
        return;
    }
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count * 16;
    if (blocks > cap) blocks = cap;
    gcn_agg_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(ih, il, ahat, oh, ol, F, C, CS, OS);
}

}  // namespace mmego
