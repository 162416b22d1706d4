// lstm_tc.cu -- IMU_Net's H=512 bidirectional LSTMs (Net/IMU_Net.py:58-62, 80, 85; 97.5 % of the pipeline's FLOPs) on
// the 5th-generation tensor cores: one persistent, warp-specialised tcgen05 kernel per timestep.
//
//   gates[M, 4H] = [x_t | h_{t-1}] [M, In+H] * W^T            M = sequences (81,920 at B=4096), 4H = 2048
//
// * CTA tile 128 sequences x 256 gate columns (= 64 hidden units x {i,f,g,o}: the weight rows are packed
//   gate-interleaved so that one tile holds all four gates of its units), K in blocks of 64.
// * warp 0 = TMA producer (operand tiles -> 128B-swizzled shared memory, mbarrier ring),
//   warp 1 = MMA issuer (tcgen05.mma, fp32 accumulators in TMEM, double buffered: 2 x 256 columns),
//   warps 2..17 = epilogue (tcgen05.ld of partial sums -> fp32 register accumulation -> bias, sigmoid/tanh, cell update,
//   h written back as the next step's operand).
// * persistent: grid = #SMs, static round-robin over (sequence tile, direction, unit tile) with the unit tile fastest,
//   so the CTAs that share an activation tile run together and hit it in L2.
// * precision: activations and weights are stored as fp16 hi/lo pairs.  NPASS = 3 evaluates
//   a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (weights pre-scaled by 2^e so that w_lo stays a normal fp16) with fp32
//   accumulation -- ~2^-22 relative, i.e. fp32-grade, which is what the 1e-3 cm parity target needs; NPASS = 1 uses the
//   hi parts only (plain fp16 tensor-core GEMM, reported with its own tolerance).
// * the recurrent operand of step t is read straight from the layer's own output tensor at t-1 (no separate h buffer);
//   the cell state is fp32, laid out [direction][unit][sequence] so that the epilogue's accesses are coalesced.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>

#include "internal.h"
#include "pack.h"
#include "tc_common.cuh"
#include "mma_frag.cuh"
#include "point_layout.h"

namespace mmego {

namespace {

using namespace tc;

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 16;
// Tile geometry as a function of the tile width BN (gate columns): BN = 256 -> 64 hidden units per tile, 2 TMEM buffers;
// BN = 128 -> 32 units per tile and FOUR TMEM buffers, so the MMA thread can run three chunks ahead of the epilogue
// warps, whose per-thread accumulator set halves (32 registers) -- no spills in the LSTM cell phase.
template <int BN>
struct Geo {
    static constexpr int UT = BN / 4;                    // hidden units per tile
    static constexpr int NT = 4 * kImuH / BN;            // unit tiles per direction
    static constexpr int UW = UT / (kEpiWarps / 4);      // units per epilogue warp (16 or 8)
    static constexpr int NBUF = 512 / BN;                // TMEM accumulator buffers
    static constexpr int W_TILE = BN * BK * 2;
};
constexpr int kThreads = 128 + kEpiWarps * 32;   // 640: warpgroup 0 = {TMA, MMA, 2 idle warps}, warpgroups 1..4 = epilogue
constexpr int A_TILE = BM * BK * 2;              // 16 KB
constexpr uint32_t kTmemCols = 512;
// Activation planes hold 2^8 * value: with |h| <= 1 (and |u| up to ~250) the hi part stays far below the fp16 maximum,
// while the lo part (residual, ~2^-12 of the value) stays a NORMAL fp16 for |value| >= 1e-3 -- unscaled, the residual of
// a typical h ~ 0.05 would be a denormal and lose most of its 11 bits.
constexpr float kActScale = 256.0f, kActInv = 1.0f / 256.0f;
// Residual ("lo") planes need far fewer than fp16's 11 significant bits: with d low mantissa bits rounded away a value
// keeps 11 + (11 - d) bits.  The tensor core's power draw depends on the operand bits (scripts/ubench/tcgen05_rate.cu:
// +5.6 % / +10.7 % MMA rate under the power cap with 5 / 8 low mantissa bits of ONE operand cleared), and the kernel is
// power-limited, so shorter residuals are faster.  Two fp16 per 32-bit word: add half an ulp of the kept precision, mask.
// (A carry out of the mantissa rounds up into the exponent, which is the correct result; residuals are never near inf.)
__host__ __device__ inline uint32_t lo_round_add(int drop) { return drop > 0 ? 0x00010001u << (drop - 1) : 0u; }
__host__ __device__ inline uint32_t lo_round_mask(int drop) { return ~(((1u << drop) - 1u) * 0x00010001u); }
__device__ __forceinline__ uint32_t lo_round(uint32_t w, uint32_t add, uint32_t mask) { return (w + add) & mask; }

// NCTA = 2: a CTA pair (cta_group::2) shares one 256 x 256 accumulator tile pair: each CTA stages its own 128 sequences
// of A and HALF of the weight tile, so a stage is 64 KB instead of 96 KB (3-deep ring instead of 2) and the weight bytes
// that cross L2 -> shared memory halve.
template <int NPASS, int NCTA, int BN>
struct Cfg {
    static constexpr int PLANES = NPASS == 3 ? 2 : 1;
    static constexpr int W_TILE_CTA = Geo<BN>::W_TILE / NCTA;
    static constexpr int STAGE_BYTES = PLANES * (A_TILE + W_TILE_CTA);      // per CTA: 96 / 48 KB (1 CTA), 64 / 32 KB (pair)
    static constexpr int STAGES = (196 * 1024) / STAGE_BYTES > 6 ? 6 : (196 * 1024) / STAGE_BYTES;
    static constexpr int BIAS_BYTES = 2 * 4 * kImuH * 4;                    // both directions' bias vectors, staged once
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + BIAS_BYTES;
};

struct StepParams {
    int S;               // sequences
    int m_tiles;
    int kb_in, kb_rec;   // K blocks of the input segment / of the recurrent segment (0 on the first step)
    int in_features;     // column of W where the recurrent block starts
    int step0, nsteps;   // this launch covers timesteps step0 .. step0+nsteps-1 (forward direction: time = step,
                         // backward: T-1-step); nsteps > 1 = persistent over timesteps, see dep_wait()
    unsigned* sync;      // nsteps > 1: arrival counters [nsteps-1][m_tiles_pad][2], zeroed before the launch
    int m_tiles_pad;     // m_tiles rounded up to a whole number of CTA groups
    unsigned* error;     // set when a bounded dependency wait gives up (mmego_debug_stats out8[7])
    int T;
    const float* bias;   // [2][2048], packed row order
    float* cstate;       // [2][512][Spad]
    long long Spad;
    __half* out_hi;      // [S][T][1024]
    __half* out_lo;      // may be null (NPASS == 1)
    float out_scale;     // 2^-e
    int kb_chunk;        // K blocks accumulated in TMEM before the partial sum is drained into registers
    int kb_chunk0;       // length of the FIRST TWO chunks of a tile: they run while the epilogue warps are still busy with
                         // the previous tile's cell update, so they are longer to give the MMA thread work until then
    unsigned long long* stats;   // dbg & 4, MMA thread: [3] work items, [4] cycles waiting for the epilogue, [5] waiting for TMA,
                         // [6] total cycles
    int dbg;             // experiment switches (results are wrong when set): 1 = skip the cell math and stores, 2 = skip the TMEM drains
    int pdl;             // launched with programmatic stream serialization: wait for the previous grid after the prologue
    uint32_t lo_add, lo_mask;   // rounding of the residual plane to fewer mantissa bits (two fp16 per word), see lo_round()
};

// Gate non-linearities straight on the SFU instructions (MUFU.EX2 / MUFU.RCP, one instruction each; the CUDA
// intrinsics wrap them in range-fixing code that triples the epilogue's instruction count).  Saturation is exact:
// ex2 -> +inf gives rcp -> 0.  Absolute error <= 2e-7 on sigmoid and tanh.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
// tanh: 1 - 2/(1 + e^2x) cancels for small |x| (absolute error ~3e-7 near 0, which dominated the error of h once the
// accumulation error was fixed), so |x| < 0.25 uses the odd Taylor polynomial up to x^9 (truncation < 2e-8 relative).
__device__ __forceinline__ float tanh_fast(float x) {
    const float x2 = x * x;
    float p = fmaf(x2, 0.021869488536155203f, -0.053968253968253968f);   // 62/2835, -17/315
    p = fmaf(x2, p, 0.13333333333333333f);                                // 2/15
    p = fmaf(x2, p, -0.33333333333333333f);                               // -1/3
    p = fmaf(x2 * x, p, x);
    const float big = fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f);
    return fabsf(x) < 0.25f ? p : big;
}
// Single-pass fp16 mode only: MUFU.TANH (relative error 2^-11, the precision h is stored with in that mode) for all five
// gate functions, sigmoid(x) = 0.5 tanh(x/2) + 0.5 -- 5 SFU instructions per cell instead of 10.
__device__ __forceinline__ float tanh_mufu(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool HALF>
__device__ __forceinline__ void lstm_cell(float pi, float pf, float pg, float po, float cprev, float& cn, float& h) {
    if (HALF) {
        const float si = fmaf(0.5f, tanh_mufu(0.5f * pi), 0.5f), sf = fmaf(0.5f, tanh_mufu(0.5f * pf), 0.5f);
        const float so = fmaf(0.5f, tanh_mufu(0.5f * po), 0.5f);
        cn = fmaf(sf, cprev, si * tanh_mufu(pg));
        h = so * tanh_mufu(cn);
    } else {
        cn = fmaf(sigmoid_fast(pf), cprev, sigmoid_fast(pi) * tanh_fast(pg));
        h = sigmoid_fast(po) * tanh_fast(cn);
    }
}

// Experiment switches (skip the cell math / the TMEM drains, cycle counters) exist only in test builds: in the product
// library they are compiled out, so no option can make the kernel produce wrong results.
#ifdef MMEGO_DEBUG_SWITCHES
#define TC_DBG(p) ((p).dbg)
#else
#define TC_DBG(p) 0
#endif

// Persistent mode (StepParams::nsteps > 1): work item (step, m tile, direction, unit tile) may read h_{step-1} of its 128
// sequences only when all unit tiles of (step-1, m tile, direction) have stored it.  Every epilogue warp of such an item
// makes its stores visible (generic -> async proxy fence, device fence) and adds 1 to the counter of (step, m, dir); the
// TMA producer of a dependent item polls the counter before its first recurrent K block.  Items are walked in (step,
// item) order by every CTA and the grid is co-resident (1 CTA per SM), so a dependency is always on an item that has
// already been started: no deadlock.  The poll is bounded and raises a device flag instead of hanging.
__device__ __forceinline__ void dep_wait(const unsigned* cnt, unsigned target, unsigned* err) {
    unsigned v = 0;
    for (int polls = 0;; ++polls) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if (v >= target) break;
        if ((polls & 255) == 255) {
            unsigned e = 0;
            if (err) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(err) : "memory");
            if (e) break;
            if (polls > (1 << 23)) {      // several seconds: only a lost dependency, never a slow neighbour, ends here
                if (err) atomicExch(err, 2u);
                break;
            }
        }
        __nanosleep(64);
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void dep_signal(unsigned* cnt) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
}

template <int NPASS, int NCTA, int BN, bool MUFU_CELL = (NPASS == 1)>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc_step_kernel(const __grid_constant__ CUtensorMap mXhi, const __grid_constant__ CUtensorMap mXlo,
                    const __grid_constant__ CUtensorMap mYhi, const __grid_constant__ CUtensorMap mYlo,
                    const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo,
                    const StepParams p) {
    using C = Cfg<NPASS, NCTA, BN>;
    using G = Geo<BN>;
    constexpr int kUnitsPerTile = G::UT, kNTiles = G::NT, kUnitsPerEpiWarp = G::UW, NBUF = G::NBUF;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the address space known to the compiler)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tfull = bars + 2 * C::STAGES;
    uint64_t* tempty = tfull + NBUF;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + NBUF);
    float* sbias = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * 4 * kImuH; i += kThreads) sbias[i] = p.bias[i];
    // work items: (sequence-tile group, direction, unit tile); a group is NCTA consecutive sequence tiles, one per CTA
    const int total_tiles = ((p.m_tiles + NCTA - 1) / NCTA) * 2 * kNTiles;
    const int total_items = total_tiles * p.nsteps;      // (step, tile), step-major
    const int kb_total = p.kb_in + p.kb_rec;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0;
    const int first_item = blockIdx.x / NCTA, item_stride = gridDim.x / NCTA;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mXhi);
        prefetch_tensormap(&mYhi);
        prefetch_tensormap(&mWhi);
        if (NPASS == 3) {
            prefetch_tensormap(&mXlo);
            prefetch_tensormap(&mYlo);
            prefetch_tensormap(&mWlo);
        }
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < NBUF; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], NCTA * kEpiWarps);      // the leader's MMA thread waits for both CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (NCTA == 2) {
            tmem_alloc_pair(tmem_holder, kTmemCols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_holder, kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2) cluster_sync_all();      // the peer's barriers must be initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    if (p.pdl) {
        // Programmatic dependent launch: everything above (barriers, TMEM, bias -- weights only) ran while the previous
        // timestep's last CTAs were still finishing; h_{t-1} and the cell state are read only after that grid has
        // completed and flushed.  The next step may start ITS prologue as soon as this grid's CTAs leave their SMs.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }

    // register re-balancing: the control warpgroup keeps 32 registers per thread (the CTA pool must balance: 128 x (96-32) = 512 x (112-96)), the 16 epilogue warps (64 fp32
    // accumulators per thread) grow to 112
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
      if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = first_item; item < total_items; item += item_stride) {
                const int sl = item / total_tiles, tile = item - sl * total_tiles, step = p.step0 + sl;
                const int nt = tile % kNTiles, dir = (tile / kNTiles) & 1, m = (tile / (2 * kNTiles)) * NCTA + (int)rank;
                const int wrow = dir * 4 * kImuH + nt * BN + (int)rank * (BN / NCTA);
                const int tt = dir ? p.T - 1 - step : step, tp = dir ? p.T - step : step - 1;
                for (int kb = 0; kb < kb_total; ++kb) {
                    if (kb == p.kb_in && sl > 0)       // h_{step-1} of these sequences: written inside this launch
                        dep_wait(p.sync + ((long long)(sl - 1) * p.m_tiles_pad + m) * 2 + dir, kNTiles * kEpiWarps, p.error);
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    uint8_t* sw = sa + C::PLANES * A_TILE;
                    int wcol, acol, at;
                    const CUtensorMap *ahi, *alo;
                    if (kb < p.kb_in) {
                        ahi = &mXhi; alo = &mXlo; acol = kb * BK; at = tt; wcol = kb * BK;
                    } else {
                        const int kr = kb - p.kb_in;
                        ahi = &mYhi; alo = &mYlo; acol = dir * kImuH + kr * BK; at = tp; wcol = p.in_features + kr * BK;
                    }
                    if (NCTA == 2) {
                        // both CTAs' bytes are counted on the LEADER's barrier; only the leader arms it
                        const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
                        if (rank == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
                        tma_load_3d_pair(sa, ahi, bar, acol, at, m * BM);
                        if (NPASS == 3) tma_load_3d_pair(sa + A_TILE, alo, bar, acol, at, m * BM);
                        tma_load_2d_pair(sw, &mWhi, bar, wcol, wrow);
                        if (NPASS == 3) tma_load_2d_pair(sw + C::W_TILE_CTA, &mWlo, bar, wcol, wrow);
                    } else {
                        mbar_expect_tx(&full[stage], C::STAGE_BYTES);
                        tma_load_3d(sa, ahi, &full[stage], acol, at, m * BM);
                        if (NPASS == 3) tma_load_3d(sa + A_TILE, alo, &full[stage], acol, at, m * BM);
                        tma_load_2d(sw, &mWhi, &full[stage], wcol, wrow);
                        if (NPASS == 3) tma_load_2d(sw + C::W_TILE_CTA, &mWlo, &full[stage], wcol, wrow);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // The tensor core's fp32 accumulator loses low-order bits on every accumulate (measured: errors grow with the
        // length of the accumulation chain and are biased, so they compound through the recurrence).  The K loop is
        // therefore cut into chunks of p.kb_chunk K-blocks: each chunk accumulates into one of the two TMEM buffers
        // from zero, and the epilogue warps drain finished chunks into fp32 registers (round-to-nearest adds) while
        // the next chunk is being multiplied.
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BM * NCTA, BN, 0 /*fp16*/);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t cc = 0;      // running chunk counter (same sequence in the epilogue warps)
            long long ms_epi = 0, ms_tma = 0, ms_tiles = 0;
            const long long ms_start = (TC_DBG(p) & 4) ? clock64() : 0;
            for (int item = first_item; item < total_items; item += item_stride) {
                if (TC_DBG(p) & 4) ++ms_tiles;
                for (int c0 = 0, ci = 0; c0 < kb_total; ++cc, ++ci) {
                    const int clen = ci < 2 ? p.kb_chunk0 : p.kb_chunk;
                    const uint32_t buf = cc % NBUF, bph = (cc / NBUF) & 1;
                    long long tm0 = 0;
                    if (TC_DBG(p) & 4) tm0 = clock64();
                    mbar_wait(&tempty[buf], bph ^ 1);
                    tc_fence_after();
                    if (TC_DBG(p) & 4) ms_epi += clock64() - tm0;
                    const uint32_t d_tmem = tmem_base + buf * BN;
                    const int c1 = min(kb_total, c0 + clen);
                    for (int kb = c0; kb < c1; ++kb) {
                        if (TC_DBG(p) & 4) tm0 = clock64();
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        if (TC_DBG(p) & 4) ms_tma += clock64() - tm0;
                        const uint32_t a_hi = smem_u32(smem + stage * C::STAGE_BYTES);
                        const uint32_t a_lo = a_hi + A_TILE;
                        const uint32_t w_hi = a_hi + C::PLANES * A_TILE;
                        const uint32_t w_lo = w_hi + C::W_TILE_CTA;
                        auto mma = [&](uint64_t da, uint64_t dw, uint32_t accumulate) {
                            if (NCTA == 2) mma_f16_ss_pair(d_tmem, da, dw, idesc, accumulate);
                            else mma_f16_ss(d_tmem, da, dw, idesc, accumulate);
                        };
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t da_hi = make_sw128_kmajor_desc(a_hi + k * 32);
                            const uint64_t dw_hi = make_sw128_kmajor_desc(w_hi + k * 32);
                            if (NPASS == 3) {
                                // small correction terms first, the large term last
                                const uint64_t da_lo = make_sw128_kmajor_desc(a_lo + k * 32);
                                const uint64_t dw_lo = make_sw128_kmajor_desc(w_lo + k * 32);
                                mma(da_hi, dw_lo, (kb > c0) || (k > 0));
                                mma(da_lo, dw_hi, 1);
                                mma(da_hi, dw_hi, 1);
                            } else {
                                mma(da_hi, dw_hi, (kb > c0) || (k > 0));
                            }
                        }
                        // frees the smem slot (in both CTAs of a pair) when these MMAs have read it
                        if (NCTA == 2) mma_commit_pair(&empty[stage], 3);
                        else mma_commit(&empty[stage]);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                    // partial accumulator complete -> epilogue warps (of both CTAs)
                    if (NCTA == 2) mma_commit_pair(&tfull[buf], 3);
                    else mma_commit(&tfull[buf]);
                    c0 = c1;
                }
            }
            if (TC_DBG(p) & 4) {
                atomicAdd(p.stats + 4, (unsigned long long)ms_epi);
                atomicAdd(p.stats + 5, (unsigned long long)ms_tma);
                atomicAdd(p.stats + 6, (unsigned long long)(clock64() - ms_start));
                atomicAdd(p.stats + 3, (unsigned long long)ms_tiles);
            }
        }
        __syncwarp();
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // ===================================================================== epilogue (16 warps)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int part = (warp - 4) >> 2;       // which 16 of the tile's 64 hidden units
        const int u0 = part * kUnitsPerEpiWarp;
        uint32_t cc = 0;
        const uint32_t tempty_remote0 = NCTA == 2 ? mapa_u32(smem_u32(&tempty[0]), 0) : 0;   // + 8 bytes per buffer
        for (int item = first_item; item < total_items; item += item_stride) {
            const int sl = item / total_tiles, tile = item - sl * total_tiles, step = p.step0 + sl;
            const int nt = tile % kNTiles, dir = (tile / kNTiles) & 1, m = (tile / (2 * kNTiles)) * NCTA + (int)rank;
            float acc[4][kUnitsPerEpiWarp];     // i, f, g, o pre-activations (scaled) of this thread's row
            const long long row = (long long)m * BM + q * 32 + lane;
            const bool ok = row < p.S;
            const bool has_state = p.kb_rec > 0;
            float* cbase = p.cstate + ((long long)(dir * kImuH + nt * kUnitsPerTile + u0)) * p.Spad + row;
            for (int c0 = 0, ci = 0; c0 < kb_total; ++cc, ++ci) {
                const int clen = ci < 2 ? p.kb_chunk0 : p.kb_chunk;
                const bool first = c0 == 0;
                const uint32_t buf = cc % NBUF, bph = (cc / NBUF) & 1;
                if (c0 + clen >= kb_total && has_state && ok) {
                    // last chunk of the tile: pull the cell state towards L2 now (no registers held), it is read below
#pragma unroll
                    for (int j = 0; j < kUnitsPerEpiWarp; ++j)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(cbase + (long long)j * p.Spad));
                }
                mbar_wait(&tfull[buf], bph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + u0);
                if (!(TC_DBG(p) & 2))
#pragma unroll
                for (int g = 0; g < 4; ++g) {        // one gate at a time keeps the temporaries small
                    uint32_t r0[kUnitsPerEpiWarp];
                    if (kUnitsPerEpiWarp == 16) tmem_ld_x16(taddr + g * kUnitsPerTile, r0);
                    else tmem_ld_x8(taddr + g * kUnitsPerTile, r0);
                    tmem_ld_wait();
                    if (first) {
#pragma unroll
                        for (int j = 0; j < kUnitsPerEpiWarp; ++j) acc[g][j] = __uint_as_float(r0[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < kUnitsPerEpiWarp; ++j) acc[g][j] += __uint_as_float(r0[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (NCTA == 2) mbar_arrive_cluster(tempty_remote0 + buf * 8);
                    else mbar_arrive(&tempty[buf]);
                }
                c0 += clen;
            }
            if (TC_DBG(p) & 1) continue;
            // ---- LSTM cell on the register-resident pre-activations (bias from shared memory: warp-wide broadcast reads)
            // The epilogue warps run at 112 registers with 64 of them holding the accumulators, and shared memory leaves almost
            // no L1, so a spilled register costs an L2 round trip: the 16 units are processed in two halves of 8 whose
            // temporaries (cell state, packed outputs) are kept small, with a scheduling fence between the halves.
            const float* bias = sbias + dir * 4 * kImuH + nt * BN + u0;
            const long long o = (row * p.T + (dir ? p.T - 1 - step : step)) * (2 * kImuH) + dir * kImuH + nt * kUnitsPerTile + u0;
#pragma unroll
            for (int half = 0; half < kUnitsPerEpiWarp / 8; ++half) {
                float cprev[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)      // 8 independent coalesced loads in flight
                    cprev[j] = (has_state && ok) ? __ldcg(cbase + (long long)(half * 8 + j) * p.Spad) : 0.f;   // L2: another SM wrote it
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) {
                    float hv2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = half * 8 + 2 * j2 + e;
                        const float pi = fmaf(acc[0][j], p.out_scale, bias[j]);
                        const float pf = fmaf(acc[1][j], p.out_scale, bias[kUnitsPerTile + j]);
                        const float pg = fmaf(acc[2][j], p.out_scale, bias[2 * kUnitsPerTile + j]);
                        const float po = fmaf(acc[3][j], p.out_scale, bias[3 * kUnitsPerTile + j]);
                        float cn, hh;
                        lstm_cell<MUFU_CELL>(pi, pf, pg, po, cprev[2 * j2 + e], cn, hh);
                        hv2[e] = hh * kActScale;
                        if (ok) __stcs(cbase + (long long)j * p.Spad, cn);
                    }
                    const __half2 hh2 = __floats2half2_rn(hv2[0], hv2[1]);           // one packed conversion
                    const float2 back = __half22float2(hh2);
                    const __half2 ll2 = __floats2half2_rn(hv2[0] - back.x, hv2[1] - back.y);
                    ph[j2] = *reinterpret_cast<const uint32_t*>(&hh2);
                    pl[j2] = lo_round(*reinterpret_cast<const uint32_t*>(&ll2), p.lo_add, p.lo_mask);
                }
                if (ok) {
                    *reinterpret_cast<uint4*>(p.out_hi + o + half * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + o + half * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                }
                asm volatile("" ::: "memory");
            }
            if (sl + 1 < p.nsteps) {
                // h_step and c of this warp's rows and units are stored: publish them to the next step's items
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __threadfence();
                __syncwarp();
                if (lane == 0) dep_signal(p.sync + ((long long)sl * p.m_tiles_pad + m) * 2 + dir);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (NCTA == 2) cluster_sync_all();      // neither CTA may exit (or free TMEM) while the other still signals it
    if (warp == 1) {
        tc_fence_after();
        if (NCTA == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------- small split-fp16 kernels
__device__ __forceinline__ void store_split(__half* hi, __half* lo, long long i, float v) {
    v = fminf(fmaxf(v * kActScale, -65000.f), 65000.f);
    const __half h = __float2half_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2half_rn(v - __half2float(h));
}

// fc1 + ReLU (Net/IMU_Net.py:79): imu [rows,15] -> u [rows,512] as fp16 hi/lo planes.  HBM-bound (writes 2 KB per
// row against 7.7 KFLOP), so the 15 -> 512 projection runs on mma.sync (mma_frag.cuh, fp16x3) to keep the issue slots
// for the split/pack/store epilogue: a warp owns 16 rows, one k-step, 64 n-tiles; the column permutation of
// pack_imu_fc1_mma gives every lane 8 consecutive channels per row = one 16-byte store per plane.
constexpr int FC1_WORDS = mma_frag_words(1, 64);
__global__ void __launch_bounds__(256, 4) imu_fc1_mma_kernel(const float* __restrict__ imu, const float* __restrict__ blob,
                                                          __half* __restrict__ uhi, __half* __restrict__ ulo,
                                                          long long rows, uint32_t lo_add, uint32_t lo_mask) {
    extern __shared__ __align__(16) uint32_t fsm[];    // frags [64][32] uint4 | bias [512]
    for (int i = threadIdx.x * 4; i < FC1_WORDS + kImuH; i += 256 * 4)
        *reinterpret_cast<uint4*>(fsm + i) = *reinterpret_cast<const uint4*>(blob + i);
    const float os = blob[FC1_WORDS + kImuH];
    __syncthreads();
    const uint4* wf = reinterpret_cast<const uint4*>(fsm);
    const float* bias = reinterpret_cast<const float*>(fsm + FC1_WORDS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
    // the 8 input values of a lane's fragment (rows g / g+8, channels 2tq, 2tq+1, 2tq+8, 2tq+9): the NEXT row tile's
    // are fetched before the current tile's 16 channel groups are computed (the loads were the kernel's main stall)
    auto load_in = [&](long long r0, float* v) {
        const long long ra = r0 + g, rb = r0 + g + 8;
        const bool la = ra < rows, lb = rb < rows;
        const float* pa = imu + ra * kImuFeat;
        const float* pb = imu + rb * kImuFeat;
        const int c = 2 * tq;
        v[0] = la ? pa[c] : 0.f; v[1] = la ? pa[c + 1] : 0.f;
        v[2] = lb ? pb[c] : 0.f; v[3] = lb ? pb[c + 1] : 0.f;
        v[4] = la ? pa[c + 8] : 0.f; v[5] = (la && c + 9 < kImuFeat) ? pa[c + 9] : 0.f;
        v[6] = lb ? pb[c + 8] : 0.f; v[7] = (lb && c + 9 < kImuFeat) ? pb[c + 9] : 0.f;
    };
    const long long rstride = (long long)gridDim.x * 128;
    float vin[8];
    {
        const long long rfirst = ((long long)blockIdx.x * 8 + warp) * 16;
        if (rfirst < rows) load_in(rfirst, vin);
    }
    for (long long r0 = ((long long)blockIdx.x * 8 + warp) * 16; r0 < rows; r0 += rstride) {
        const long long ra = r0 + g, rb = r0 + g + 8;
        const bool la = ra < rows, lb = rb < rows;
        uint32_t ah[1][4], al[1][4];
        frag::split2(vin[0], vin[1], ah[0][0], al[0][0]);
        frag::split2(vin[2], vin[3], ah[0][1], al[0][1]);
        frag::split2(vin[4], vin[5], ah[0][2], al[0][2]);
        frag::split2(vin[6], vin[7], ah[0][3], al[0][3]);
        if (r0 + rstride < rows) load_in(r0 + rstride, vin);
#pragma unroll 2
        for (int q = 0; q < 16; ++q) {
            float out[4][4];
            frag::dense_tile<1, 4, true, 64>(wf, bias, os, ah, al, out, lane, 4 * q);
            uint32_t h0[4], l0[4], h1[4], l1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                frag::split2(fminf(out[j][0] * kActScale, 65000.f), fminf(out[j][1] * kActScale, 65000.f), h0[j], l0[j]);
                frag::split2(fminf(out[j][2] * kActScale, 65000.f), fminf(out[j][3] * kActScale, 65000.f), h1[j], l1[j]);
                l0[j] = lo_round(l0[j], lo_add, lo_mask);
                l1[j] = lo_round(l1[j], lo_add, lo_mask);
            }
            const long long ca = ra * kImuH + 32 * q + 8 * tq, cb = rb * kImuH + 32 * q + 8 * tq;
            if (la) {
                *reinterpret_cast<uint4*>(uhi + ca) = make_uint4(h0[0], h0[1], h0[2], h0[3]);
                if (ulo) *reinterpret_cast<uint4*>(ulo + ca) = make_uint4(l0[0], l0[1], l0[2], l0[3]);
            }
            if (lb) {
                *reinterpret_cast<uint4*>(uhi + cb) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
                if (ulo) *reinterpret_cast<uint4*>(ulo + cb) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
            }
        }
    }
}

__device__ __forceinline__ void load8(const __half* hi, const __half* lo, long long i, float* v) {
    const uint4 a = *reinterpret_cast<const uint4*>(hi + i);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(ah[k]);
        v[2 * k] = f.x;
        v[2 * k + 1] = f.y;
    }
    if (lo) {
        const uint4 b = *reinterpret_cast<const uint4*>(lo + i);
        const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(bh[k]);
            v[2 * k] += f.x;
            v[2 * k + 1] += f.y;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= kActInv;
}

// attention pooling over the n samples of a frame (Net/IMU_Net.py:82-83) on split planes: y [F,n,1024] -> s [F,1024].
// HBM-bound: 4 KB (8 KB with the lo plane) in per sample, read ONCE.  One warp owns a frame: lane l holds channels
// i*256 + 8l .. +7 (i = 0..3) of every sample, the score is a warp reduction, and the softmax is folded in online
// (running max / sum / weighted channel sums), so there is no second pass over y, no shared memory and no block barrier.
// The next sample's 8 x 16 B loads are issued before the current one is consumed.
struct PoolRow {
    uint4 h[4], l[4];
};
__device__ __forceinline__ void pool_load(const __half* yhi, const __half* ylo, long long off, int lane, PoolRow& r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.h[i] = *reinterpret_cast<const uint4*>(yhi + off + i * 256 + lane * 8);
        if (ylo) r.l[i] = *reinterpret_cast<const uint4*>(ylo + off + i * 256 + lane * 8);
    }
}
__device__ __forceinline__ void pool_unpack(const PoolRow& r, bool has_lo, float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2* ah = reinterpret_cast<const __half2*>(&r.h[i]);
        const __half2* bh = reinterpret_cast<const __half2*>(&r.l[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 f = __half22float2(ah[k]);
            if (has_lo) {
                const float2 g = __half22float2(bh[k]);
                f.x += g.x;
                f.y += g.y;
            }
            v[i * 8 + 2 * k] = f.x * kActInv;
            v[i * 8 + 2 * k + 1] = f.y * kActInv;
        }
    }
}
constexpr int POOL_WARPS = 4;
__global__ void __launch_bounds__(POOL_WARPS * 32) imu_pool_split_kernel(const __half* __restrict__ yhi,
                                                                        const __half* __restrict__ ylo,
                                                                        const float* __restrict__ attn,
                                                                        __half* __restrict__ shi, __half* __restrict__ slo,
                                                                        long long F, int n, uint32_t lo_add, uint32_t lo_mask) {
    const int lane = threadIdx.x & 31;
    const long long f = (long long)blockIdx.x * POOL_WARPS + (threadIdx.x >> 5);
    if (f >= F) return;
    float aw[32];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) aw[i * 8 + k] = attn[i * 256 + lane * 8 + k];
    const float ab = attn[1024];
    const bool has_lo = ylo != nullptr;
    const long long base = f * (long long)n * 1024;
    float m = -INFINITY, sum = 0.f, acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    PoolRow cur, nxt;
    pool_load(yhi, ylo, base, lane, cur);
    for (int s0 = 0; s0 < n; ++s0) {
        if (s0 + 1 < n) pool_load(yhi, ylo, base + (long long)(s0 + 1) * 1024, lane, nxt);
        float v[32];
        pool_unpack(cur, has_lo, v);
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) a = fmaf(v[k], aw[k], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        a += ab;
        const float nm = fmaxf(m, a);
        const float r = expf(m - nm);          // first sample: exp(-inf) = 0
        const float w = expf(a - nm);
        sum = fmaf(sum, r, w);
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = fmaf(acc[k], r, w * v[k]);
        m = nm;
        cur = nxt;
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float val = fminf(fmaxf(acc[i * 8 + k] * inv * kActScale, -65000.f), 65000.f);
            hi[k] = __float2half_rn(val);
            lo[k] = __float2half_rn(val - __half2float(hi[k]));
        }
        const long long o = f * 1024 + i * 256 + lane * 8;
        *reinterpret_cast<uint4*>(shi + o) = *reinterpret_cast<const uint4*>(hi);
        if (slo) {
            uint4 lw = *reinterpret_cast<const uint4*>(lo);
            lw.x = lo_round(lw.x, lo_add, lo_mask); lw.y = lo_round(lw.y, lo_add, lo_mask);
            lw.z = lo_round(lw.z, lo_add, lo_mask); lw.w = lo_round(lw.w, lo_add, lo_mask);
            *reinterpret_cast<uint4*>(slo + o) = lw;
        }
    }
}

__device__ __forceinline__ void ortho6d_cols(const float* a6, float eps, float* m) {
    float ax = a6[0], ay = a6[1], az = a6[2];
    const float bx = a6[3], by = a6[4], bz = a6[5];
    float n = fmaxf(sqrtf(ax * ax + ay * ay + az * az), eps);
    ax /= n; ay /= n; az /= n;
    float zx = ay * bz - az * by, zy = az * bx - ax * bz, zz = ax * by - ay * bx;
    n = fmaxf(sqrtf(zx * zx + zy * zy + zz * zz), eps);
    zx /= n; zy /= n; zz /= n;
    const float yx = zy * az - zz * ay, yy = zz * ax - zx * az, yz = zx * ay - zy * ax;
    m[0] = ax; m[1] = yx; m[2] = zx;
    m[3] = ay; m[4] = yy; m[5] = zy;
    m[6] = az; m[7] = yz; m[8] = zz;
}

// fc2 + ortho6d (Net/IMU_Net.py:87-93) on split planes; one warp per frame, persistent CTAs with fc2 (36 KB) staged
// once in shared memory (reading it through L1/L2 for every frame made this kernel L2-bound at 9x its HBM bytes).
// Lane l holds channels i*128 + 4l .. +3 (i = 0..7): 8-byte plane loads, conflict-free 16-byte weight reads.
__global__ void __launch_bounds__(256) imu_decode_split_kernel(const __half* __restrict__ ghi, const __half* __restrict__ glo,
                                                               const float* __restrict__ fc2, float* __restrict__ R,
                                                               float* __restrict__ t, long long F) {
    extern __shared__ __align__(16) float sw[];        // [9][1024] + [9]
    for (int i = threadIdx.x * 4; i < 9 * 1024; i += 256 * 4)
        *reinterpret_cast<float4*>(sw + i) = *reinterpret_cast<const float4*>(fc2 + i);
    if (threadIdx.x < 9) sw[9 * 1024 + threadIdx.x] = fc2[9 * 1024 + threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // two frames per warp and iteration: every 16-byte weight read from shared memory feeds both (the kernel was bound by
    // shared-memory bandwidth: 36 KB of weights per 4 KB frame)
    for (long long f0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2; f0 < F; f0 += (long long)gridDim.x * 16) {
        const int nf = f0 + 1 < F ? 2 : 1;
        float x[2][32];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const long long f = e < nf ? f0 + e : f0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint2 a = *reinterpret_cast<const uint2*>(ghi + f * 1024 + i * 128 + lane * 4);
                float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
                float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
                if (glo) {
                    const uint2 b = *reinterpret_cast<const uint2*>(glo + f * 1024 + i * 128 + lane * 4);
                    const float2 q0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x));
                    const float2 q1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
                    p0.x += q0.x; p0.y += q0.y; p1.x += q1.x; p1.y += q1.y;
                }
                x[e][i * 4] = p0.x * kActInv; x[e][i * 4 + 1] = p0.y * kActInv;
                x[e][i * 4 + 2] = p1.x * kActInv; x[e][i * 4 + 3] = p1.y * kActInv;
            }
        }
        float T9[2][9];
#pragma unroll
        for (int o = 0; o < 9; ++o) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 wv = *reinterpret_cast<const float4*>(sw + o * 1024 + i * 128 + lane * 4);
                a0 = fmaf(wv.x, x[0][i * 4], a0);
                a0 = fmaf(wv.y, x[0][i * 4 + 1], a0);
                a0 = fmaf(wv.z, x[0][i * 4 + 2], a0);
                a0 = fmaf(wv.w, x[0][i * 4 + 3], a0);
                a1 = fmaf(wv.x, x[1][i * 4], a1);
                a1 = fmaf(wv.y, x[1][i * 4 + 1], a1);
                a1 = fmaf(wv.z, x[1][i * 4 + 2], a1);
                a1 = fmaf(wv.w, x[1][i * 4 + 3], a1);
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, sft);
                a1 += __shfl_xor_sync(0xffffffffu, a1, sft);
            }
            T9[0][o] = a0 + sw[9 * 1024 + o];
            T9[1][o] = a1 + sw[9 * 1024 + o];
        }
        if (lane < nf) {
            const long long f = f0 + lane;
            float tt[9], mm[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) tt[k] = lane == 0 ? T9[0][k] : T9[1][k];
            ortho6d_cols(tt, 1e-8f, mm);
#pragma unroll
            for (int k = 0; k < 9; ++k) R[f * 9 + k] = mm[k];
            t[f * 3] = tt[6]; t[f * 3 + 1] = tt[7]; t[f * 3 + 2] = tt[8];
        }
    }
}

// test hook: split planes -> fp32
__global__ void unsplit_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out,
                               long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (__half2float(hi[i]) + (lo ? __half2float(lo[i]) : 0.f)) * kActInv;
}

// rounds a packed residual plane (weights) in place, see lo_round()
__global__ void lo_round_kernel(uint32_t* __restrict__ w, long long nwords, uint32_t add, uint32_t mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) w[i] = lo_round(w[i], add, mask);
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

// activations [S][T][C] fp16; box = 64 channels x 1 step x 128 sequences
bool make_act_map(CUtensorMap* m, const void* base, long long S, int T, int C) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
    cuuint32_t box[3] = {BK, 1, BM};
    cuuint32_t es[3] = {1, 1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weights [rows][K] fp16; box = 64 x box_rows (256 for a single CTA, 128 = one CTA's half for a CTA pair)
bool make_w_map(CUtensorMap* m, const void* base, int rows, int K, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// ================================================================================================ host interface
bool tc_supported() { return encode_fn() != nullptr; }

// Packs one bidirectional H=512 layer for a tile of UT hidden units (64 for BN = 256, 32 for BN = 128): rows are
// gate-interleaved per tile,
//   packed row p = dir*2048 + tile*4UT + gate*UT + e   <->   torch row gate*512 + tile*UT + e   (gates i,f,g,o),
// columns [W_ih (In) | W_hh (512)], scaled by 2^e and split into fp16 hi/lo.
static bool pack_variant(mmego_handle* h, const StateDict& sd, const std::string& prefix, int layer, int In, int UT,
                         TcLstmVariant& out, float* out_scale) {
    const int H = kImuH, K = In + H, rows = 2 * 4 * H, BNv = 4 * UT;
    std::vector<float> w((size_t)rows * K), bias(rows);
    const char* sfx[2] = {"", "_reverse"};
    float wmax = 0.f;
    for (int d = 0; d < 2; ++d) {
        const std::string k = "l" + std::to_string(layer) + sfx[d];
        const float* wih = sd.get(prefix + "weight_ih_" + k, (long long)4 * H * In);
        const float* whh = sd.get(prefix + "weight_hh_" + k, (long long)4 * H * H);
        const float* bih = sd.get(prefix + "bias_ih_" + k, 4 * H);
        const float* bhh = sd.get(prefix + "bias_hh_" + k, 4 * H);
        for (int pr = 0; pr < 4 * H; ++pr) {
            const int tile = pr / BNv, gate = (pr % BNv) / UT, e = pr % UT;
            const int r = gate * H + tile * UT + e;
            float* dst = &w[((size_t)d * 4 * H + pr) * K];
            std::memcpy(dst, wih + (size_t)r * In, sizeof(float) * In);
            std::memcpy(dst + In, whh + (size_t)r * H, sizeof(float) * H);
            bias[(size_t)d * 4 * H + pr] = bih[r] + bhh[r];
        }
    }
    for (float v : w) wmax = std::max(wmax, std::fabs(v));
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
    if (!up(hi.data(), hi.size() * 2, &dhi) || !up(lo.data(), lo.size() * 2, &dlo) || !up(bias.data(), bias.size() * 4, &db))
        return false;
    out.whi = dhi;
    out.wlo = dlo;
    out.bias = static_cast<float*>(db);
    *out_scale = kActInv / scale;
    static_assert(sizeof(CUtensorMap) == sizeof(out.map_hi), "tensor map storage size");
    return make_w_map(reinterpret_cast<CUtensorMap*>(&out.map_hi), dhi, rows, K, BNv) &&
           make_w_map(reinterpret_cast<CUtensorMap*>(&out.map_lo), dlo, rows, K, BNv) &&
           make_w_map(reinterpret_cast<CUtensorMap*>(&out.map_hi2), dhi, rows, K, BNv / 2) &&
           make_w_map(reinterpret_cast<CUtensorMap*>(&out.map_lo2), dlo, rows, K, BNv / 2);
}

bool tc_pack_layer(mmego_handle* h, const StateDict& sd, const std::string& prefix, int layer, int In, TcLstmLayer& out) {
    out.in_features = In;
    out.K = In + kImuH;
    // Only the 256-column layout is packed: the 128-column / four-TMEM-buffer variant of the kernel (Geo<128>) was
    // measured at 99 ms against 72 ms for the rnn_fast launches of a B=4096 step -- every activation tile is then
    // fetched by 16 CTAs instead of 8 and the operand traffic, not the epilogue, becomes the limit.
    return pack_variant(h, sd, prefix, layer, In, 64, out.v[0], &out.out_scale);
}

template <int NPASS, int NCTA, int BN, bool MUFU_CELL = (NPASS == 1)>
cudaError_t launch_step(int grid, cudaStream_t st, const CUtensorMap& mXhi, const CUtensorMap& mXlo, const CUtensorMap& mYhi,
                        const CUtensorMap& mYlo, const CUtensorMap& mWhi, const CUtensorMap& mWlo, const StepParams& p) {
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set))
        cudaFuncSetAttribute(lstm_tc_step_kernel<NPASS, NCTA, BN, MUFU_CELL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg<NPASS, NCTA, BN>::SMEM_BYTES);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg<NPASS, NCTA, BN>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    int na = 1;
    if (p.nsteps > 1) {
        // items of this launch wait for each other: the grid must be co-resident as a whole.  A cooperative launch is
        // placed all-or-nothing, so another stream's (or process's) kernels can never hold the SMs part of it needs.
        attr[na].id = cudaLaunchAttributeCooperative;
        attr[na].val.cooperative = 1;
        ++na;
    } else if (p.pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, lstm_tc_step_kernel<NPASS, NCTA, BN, MUFU_CELL>, mXhi, mXlo, mYhi, mYlo, mWhi, mWlo, p);
}

// One bidirectional layer: x planes [S][T][In] -> y planes [S][T][1024]; cstate [2][512][Spad] fp32 scratch.
int tc_lstm_layer(mmego_handle* h, const TcLstmLayer& lw, const void* xhi, const void* xlo, void* yhi, void* ylo,
                  float* cstate, long long S, long long Spad, int T, int npass, bool persist_steps, cudaStream_t st) {
    CUtensorMap mXhi, mXlo, mYhi, mYlo;
    const int In = lw.in_features;
    if (!make_act_map(&mXhi, xhi, S, T, In) || !make_act_map(&mYhi, yhi, S, T, 2 * kImuH)) return -1;
    if (npass == 3) {
        if (!make_act_map(&mXlo, xlo, S, T, In) || !make_act_map(&mYlo, ylo, S, T, 2 * kImuH)) return -1;
    } else {
        mXlo = mXhi;
        mYlo = mYhi;
    }
    const bool pair = h->tc_cta_pair != 0;
    const int bn = 256;
    const TcLstmVariant& wv = lw.v[0];
    const CUtensorMap& mWhi = *reinterpret_cast<const CUtensorMap*>(pair ? &wv.map_hi2 : &wv.map_hi);
    const CUtensorMap& mWlo = *reinterpret_cast<const CUtensorMap*>(pair ? &wv.map_lo2 : &wv.map_lo);
    StepParams p{};
    p.S = (int)S;
    p.m_tiles = (int)((S + BM - 1) / BM);
    p.kb_in = In / BK;
    p.in_features = In;
    p.T = T;
    p.bias = wv.bias;
    p.cstate = cstate;
    p.Spad = Spad;
    p.out_hi = static_cast<__half*>(yhi);
    p.out_lo = npass == 3 ? static_cast<__half*>(ylo) : nullptr;
    p.out_scale = lw.out_scale;
    const int chunk_opt = h->tc_kb_chunk;
    p.dbg = h->tc_dbg;
    p.pdl = h->tc_pdl;
    p.lo_add = lo_round_add(h->tc_lo_drop);
    p.lo_mask = lo_round_mask(h->tc_lo_drop);
    if ((p.dbg & 4) && !h->tc_stats) {
        cudaMalloc(&h->tc_stats, 8 * sizeof(unsigned long long));
        cudaMemset(h->tc_stats, 0, 8 * sizeof(unsigned long long));
    }
    p.stats = static_cast<unsigned long long*>(h->tc_stats);
    const int ncta = pair ? 2 : 1;
    const int total = ((p.m_tiles + ncta - 1) / ncta) * 2 * (4 * kImuH / bn);  // work items (one per CTA or CTA pair)
    const int slots = h->sm_count / ncta;     // co-resident CTAs (pairs): the kernel takes a whole SM
    p.m_tiles_pad = ((p.m_tiles + ncta - 1) / ncta) * ncta;
    if (!h->dev_error) {
        if (cudaMalloc(reinterpret_cast<void**>(&h->dev_error), 64) != cudaSuccess) return -1;
        cudaMemsetAsync(h->dev_error, 0, 64, st);
    }
    p.error = h->dev_error;
    // Steps 1 .. T-1 have the same shape (step 0 has no recurrent K blocks): with `tc_persist` they run as ONE launch
    // whose work items carry their timestep and wait for the h_{t-1} they read (dep_wait) -- no per-step tail where
    // the last round of items leaves most CTA pairs idle (rnn_slow: 256 items on 74 pairs = 3.46 rounds paid as 4).
    const bool persist = persist_steps && T > 2 && !(p.dbg & 3);
    if (persist) {
        const size_t words = (size_t)(T - 2) * p.m_tiles_pad * 2;
        if (h->tc_sync_words < words) {
            if (h->tc_sync) cudaFree(h->tc_sync);
            h->tc_sync = nullptr;
            h->tc_sync_words = 0;
            if (cudaMalloc(reinterpret_cast<void**>(&h->tc_sync), words * sizeof(unsigned)) != cudaSuccess) return -1;
            h->tc_sync_words = words;
        }
        cudaMemsetAsync(h->tc_sync, 0, words * sizeof(unsigned), st);
        p.sync = h->tc_sync;
    }
    bool use_persist = persist;
    for (int step = 0; step < T;) {
        const int nsteps = (use_persist && step > 0) ? T - step : 1;
        p.step0 = step;
        p.nsteps = nsteps;
        p.pdl = nsteps > 1 ? 0 : h->tc_pdl;      // the persistent launch is cooperative, not programmatic
        p.kb_rec = step > 0 ? kImuH / BK : 0;
        p.kb_chunk = (npass == 3 && chunk_opt > 0) ? chunk_opt : (p.kb_in + p.kb_rec);
        // with four TMEM buffers (BN = 128) the MMA thread has enough run-ahead without longer first chunks
        p.kb_chunk0 = (npass == 3 && chunk_opt > 0) ? (bn == 128 ? chunk_opt : std::max(chunk_opt, h->tc_kb_chunk0)) : p.kb_chunk;
        // a persistent launch may use every slot even when one step has fewer items: items of later steps start on the
        // idle pairs, multiply their input K blocks and wait (in the TMA producer) only for the recurrent ones
        const long long items = (long long)total * nsteps;
        const int grid = (int)(items < slots ? items : slots) * ncta;
        cudaError_t le;
#define MMEGO_STEP(NP, NC, BNV) le = launch_step<NP, NC, BNV>(grid, st, mXhi, mXlo, mYhi, mYlo, mWhi, mWlo, p)
        if (npass == 3) { if (pair) MMEGO_STEP(3, 2, 256); else MMEGO_STEP(3, 1, 256); }
        else { if (pair) MMEGO_STEP(1, 2, 256); else MMEGO_STEP(1, 1, 256); }
#undef MMEGO_STEP
        if (le != cudaSuccess) {
            cudaGetLastError();
            if (nsteps > 1) {          // the device cannot place the whole grid at once: one launch per timestep instead
                use_persist = false;
                continue;
            }
            return -1;
        }
        ++t_launches;
        step += nsteps;
    }
    return 0;
}

void tc_round_lo_weights(const TcLstmLayer& lw, int drop, cudaStream_t st) {
    const long long nwords = (long long)2 * 4 * kImuH * lw.K / 2;
    if (drop <= 0 || !lw.v[0].wlo) return;
    lo_round_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(static_cast<uint32_t*>(lw.v[0].wlo), nwords,
                                                                      lo_round_add(drop), lo_round_mask(drop));
}

void tc_imu_fc1(const float* imu, const float* fc1_mma, void* uhi, void* ulo, long long rows, int sm_count, int lo_drop,
                cudaStream_t st) {
    if (rows <= 0) return;
    const int smem = (FC1_WORDS + kImuH) * 4;
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set))
        cudaFuncSetAttribute(imu_fc1_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long blocks = (rows + 127) / 128;
    if (blocks > 4LL * sm_count) blocks = 4LL * sm_count;
    ++t_launches;
    imu_fc1_mma_kernel<<<(unsigned)blocks, 256, smem, st>>>(imu, fc1_mma, static_cast<__half*>(uhi),
                                                            static_cast<__half*>(ulo), rows, lo_round_add(lo_drop),
                                                            lo_round_mask(lo_drop));
}
void tc_imu_pool(const void* yhi, const void* ylo, const float* attn, void* shi, void* slo, long long F, int n,
                 int lo_drop, cudaStream_t st) {
    if (F <= 0) return;
    ++t_launches;
    imu_pool_split_kernel<<<(unsigned)((F + POOL_WARPS - 1) / POOL_WARPS), POOL_WARPS * 32, 0, st>>>(
        static_cast<const __half*>(yhi), static_cast<const __half*>(ylo), attn, static_cast<__half*>(shi),
        static_cast<__half*>(slo), F, n, lo_round_add(lo_drop), lo_round_mask(lo_drop));
}
void tc_imu_decode(const void* ghi, const void* glo, const float* fc2, float* R, float* t, long long F, cudaStream_t st) {
    if (F <= 0) return;
    const int smem = (9 * 1024 + 16) * (int)sizeof(float);
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set))
        cudaFuncSetAttribute(imu_decode_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (F + 15) / 16;
    if (blocks > 2LL * sms) blocks = 2LL * sms;
    ++t_launches;
    imu_decode_split_kernel<<<(unsigned)blocks, 256, smem, st>>>(static_cast<const __half*>(ghi),
                                                                 static_cast<const __half*>(glo), fc2, R, t, F);
}
void tc_unsplit(const void* hi, const void* lo, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return;
    ++t_launches;
    unsplit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const __half*>(hi),
                                                                static_cast<const __half*>(lo), out, n);
}

}  // namespace mmego
