// lstm_resident.cu -- the LATENCY path of IMU_Net (Net/IMU_Net.py:67-94) for small batches: the reference's own operating
// point is ONE snippet per call (Processor/Test/Demo_test.py:61), i.e. 20 sequences for rnn_fast and a single sequence
// for rnn_slow.  The tcgen05 kernel of lstm_tc.cu is built for 40,960 sequences per launch; at 20 it spends 80 timestep
// launches streaming the 92 MB of weights through 16 CTA pairs (2.1 ms per snippet).
//
// Here one PERSISTENT kernel per bidirectional layer keeps the gate weights resident in shared memory for all
// timesteps.  Two forms of the per-step product share the protocol below: exact fp32 FMAs (described first; used for
// layers with <= 4 sequences, i.e. rnn_slow, where the product is negligible next to the inter-CTA exchange) and mma.sync
// with fp16 hi/lo split operands (segment_tc / run_steps_tc; rnn_fast's 20 sequences per snippet, where the fp32 form
// is bound by the shared-memory pipe: 11.9 -> 6.7 us per step):
//   * 128 CTAs = 2 directions x 64 groups of 8 hidden units; a CTA owns the 32 gate rows (i,f,g,o of its 8 units) over
//     the full K = In + 512: 128 / 192 KB of shared memory, loaded once per launch.
//   * per timestep a warp's lanes are the 32 gate rows, warps split the (sequence group, K slice) items; partial sums
//     meet in shared memory, then one thread per (sequence, unit) applies the cell update (c in a small global scratch).
//   * the 64 CTAs of a direction exchange h through the layer's own output tensor in global memory (L2): a CTA publishes
//     its 8 units with a device-scope fence + one arrival on a per-(direction, step) counter and waits for 64 arrivals
//     before the recurrent half of the next step.  The INPUT half of a step (x_t W_ih^T, K = In) does not depend on the
//     other CTAs and runs before the wait, hiding the exchange latency.
//   * launched with cudaLaunchCooperativeKernel (all CTAs co-resident, or the launch fails); the wait gives up after a
//     bounded number of polls and raises a flag instead of hanging the GPU.
// Everything around it (fc1, attention pooling, fc2 + 6D decode) is fp32; with imu_res_tc = 0 the whole path is the
// exact-fp32 evaluation of IMU_Net.  The host picks this path when B*L <= imu_res_max_seq (kResMaxSeq).
#include "internal.h"
#include "mma_frag.cuh"
#include "pack.h"
#ifndef MMEGO_EMUL
#include <cuda_pipeline.h>
#endif

namespace mmego {

namespace {

constexpr int RU = 8;                    // hidden units per CTA
constexpr int RR = 4 * RU;               // gate rows per CTA = lanes of a warp
constexpr int RGROUPS = kImuH / RU;      // 64 CTAs per direction
constexpr int RT = 256, RW = RT / 32;    // threads / warps per CTA
constexpr int SBLK = 20;                 // sequences per block of the K loop (4 sequence groups x 5)

#ifdef MMEGO_EMUL
#define RES_LDCG(p) (*(p))
#else
#define RES_LDCG(p) __ldcg(p)
#endif

struct ResParams {
    const float* x;        // [S][T][In]
    float* y;              // [S][T][1024]  (fwd -> 0..511, bwd -> 512..1023); also the recurrent operand
    const float* w;        // [2][64][K][32]  k-major slices, col = gate*8 + unit; tensor-core form: [2][64][K/16][4][32] uint4 B fragments
    const float* wscale;   // tensor-core form: [2][64] output scale of a slice's fragments (pack_mma_weight), else null
    const float* bias;     // [2][64][32]     b_ih + b_hh
    float* cstate;         // [2][S][512]
    unsigned* flags;       // [2][T] arrival counters, zero at launch (null: no inter-CTA waiting, one step per launch)
    unsigned* error;       // set to 1 when a wait gave up
    unsigned long long* xchg;   // [2 dirs][2 step parities][S][512] tagged h words (tag << 32 | float bits), zero at launch, or null:
                           // the exchange of h between the 64 CTAs of a direction without fence / arrival counter / poll, see wait_tagged
    int direct;            // tensor-core form: A fragments straight from global memory (segment_tc_direct) instead of the staged chunks
    float* gxs;            // [128 CTAs][S][T][32] input projections of ALL timesteps (S <= 4 only, else null), see precompute_inputs
    int S, T, In, K;
    int t_begin, t_end;
};

__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }

// ---- exchange of h_t between the CTAs of a direction through self-validating words ---------------------------------
// The first protocol (still used when p.xchg is null) publishes h in the output tensor, fences, adds one arrival per CTA
// to a per-step counter, and the readers poll the counter, fence, and only then fetch h from L2: four dependent L2 round
// trips per timestep.  Here every h value travels as ONE naturally aligned 64-bit word (step tag << 32 | float bits),
// written with st.relaxed.gpu and read with ld.relaxed.gpu: a 64-bit scalar access is single-copy atomic, so a reader
// that sees the expected tag has the value of that step -- no fence, no counter, and the wait IS the load.  Two buffers
// by step parity: a CTA writes step t+1 only after it has read step t from everybody, i.e. after every reader of step
// t-1 (same parity) is done.  The buffer is zeroed by the launch and tags start at 1.
__device__ __forceinline__ unsigned long long ld_tagged(const unsigned long long* ptr) {
#ifdef MMEGO_EMUL
    return *ptr;
#else
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_tagged(unsigned long long* ptr, unsigned tag, float v) {
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
#ifdef MMEGO_EMUL
    *ptr = w;
#else
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ptr), "l"(w) : "memory");
#endif
}
// re-reads the word until it carries `tag` (bounded: a time-out raises the error flag instead of hanging the GPU)
__device__ __forceinline__ float wait_tagged(const unsigned long long* ptr, unsigned long long w, unsigned tag, unsigned* error) {
    unsigned polls = 0;
    while ((unsigned)(w >> 32) != tag) {
        if (++polls > (1u << 22)) {
            if (error) *error = 1u;
            break;
        }
        w = ld_tagged(ptr);
    }
    return __uint_as_float((unsigned)w);
}

// One K segment (the input half x_t or the recurrent half h_{t-1}) of one block of <= 4 SQ sequences:
//   acc[r][j] += sum_k W[kw0 + k][4 rg + r] * a_{sg SQ + j}[k],  k in [0, len),  a_s = src + s * stride.
// The activations are streamed through shared memory in chunks of KC = 128 k with cp.async (16 bytes per copy, rows kept
// as they are in global memory: [sequence][k], row stride KC + 4 floats so that the four sequence groups of a warp hit
// different banks), three buffers, two chunks in flight while one is consumed -- h comes from L2 (other CTAs wrote it)
// and that latency must not sit inside the FMA loop.  Within a chunk warp w owns k-quads w, w + 8, ...; a lane is (row
// group rg = lane / 4, sequence group sg = lane % 4) with a 4 x SQ register tile: per k-quad 4 weight reads + SQ
// activation reads (all 16-byte, conflict-free) for 16 SQ FMAs.
// (Measured dead ends, B200: one gate row per lane with broadcast activation reads -- 21 shared-memory wavefronts per
// k, 30 FMA/clk/SM; activations straight from global memory into a register pipeline -- 4 sectors per load instruction,
// L1/TEX-bound; register-staged chunks with one chunk in flight -- the loop waited on L2.)
constexpr int KC = 128, KROW = KC + 4, NBUF = 3;
constexpr int ABUF = NBUF * SBLK * KROW;     // floats (31.7 KB); afterwards the cross-warp partial sums [warp][sequence][row] (20 KB)

__device__ __forceinline__ void copy16_async(float* dst, const float* src) {
#ifdef MMEGO_EMUL
    *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
#else
    __pipeline_memcpy_async(dst, src, 16);       // cp.async.cg: global -> shared without a register round trip, L2 only
#endif
}
__device__ __forceinline__ void copy_commit() {
#ifndef MMEGO_EMUL
    __pipeline_commit();
#endif
}
template <int N>
__device__ __forceinline__ void copy_wait_prior() {
#ifndef MMEGO_EMUL
    __pipeline_wait_prior(N);
#endif
}

// the FMAs of one staged chunk: buf = the lane's first sequence row of the chunk, wc = the lane's 4 weight columns at the chunk's first k
template <int SQ>
__device__ __forceinline__ void chunk_fma(const float* __restrict__ wc, const float* __restrict__ buf, int warp, float (&acc)[4][SQ]) {
#pragma unroll
    for (int q = warp; q < KC / 4; q += RW) {
        const float* wq = wc + (size_t)q * 4 * RR;
        const float4 w0 = *reinterpret_cast<const float4*>(wq), w1 = *reinterpret_cast<const float4*>(wq + RR);
        const float4 w2 = *reinterpret_cast<const float4*>(wq + 2 * RR), w3 = *reinterpret_cast<const float4*>(wq + 3 * RR);
#pragma unroll
        for (int j = 0; j < SQ; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(buf + j * KROW + 4 * q);
            acc[0][j] = fmaf(w0.x, v.x, acc[0][j]); acc[1][j] = fmaf(w0.y, v.x, acc[1][j]);
            acc[2][j] = fmaf(w0.z, v.x, acc[2][j]); acc[3][j] = fmaf(w0.w, v.x, acc[3][j]);
            acc[0][j] = fmaf(w1.x, v.y, acc[0][j]); acc[1][j] = fmaf(w1.y, v.y, acc[1][j]);
            acc[2][j] = fmaf(w1.z, v.y, acc[2][j]); acc[3][j] = fmaf(w1.w, v.y, acc[3][j]);
            acc[0][j] = fmaf(w2.x, v.z, acc[0][j]); acc[1][j] = fmaf(w2.y, v.z, acc[1][j]);
            acc[2][j] = fmaf(w2.z, v.z, acc[2][j]); acc[3][j] = fmaf(w2.w, v.z, acc[3][j]);
            acc[0][j] = fmaf(w3.x, v.w, acc[0][j]); acc[1][j] = fmaf(w3.y, v.w, acc[1][j]);
            acc[2][j] = fmaf(w3.z, v.w, acc[2][j]); acc[3][j] = fmaf(w3.w, v.w, acc[3][j]);
        }
    }
}

template <int SQ>
__device__ __forceinline__ void segment(const float* __restrict__ sw, float* __restrict__ abuf, const float* __restrict__ src,
                                        size_t stride, int nseq, int kw0, int len, int tid, float (&acc)[4][SQ]) {
    constexpr int SP = 4 * SQ;                         // sequence slots per block
    const int lane = tid & 31, warp = tid >> 5, rg = lane >> 2, sg = lane & 3;
    const int nchunk = len / KC;                       // len is a multiple of 128
    auto issue = [&](int c) {
        if (c < nchunk) {
            float* buf = abuf + (c % NBUF) * (SBLK * KROW);
            for (int i = tid; i < SP * (KC / 4); i += RT) {
                const int sq = i / (KC / 4), kq = i % (KC / 4);
                copy16_async(buf + sq * KROW + 4 * kq, src + (size_t)(sq < nseq ? sq : 0) * stride + c * KC + 4 * kq);
            }
        }
        copy_commit();                                 // (an empty group keeps the wait counts uniform)
    };
    issue(0);
    issue(1);
    for (int c = 0; c < nchunk; ++c) {
        issue(c + 2);
        copy_wait_prior<2>();                          // chunk c has landed (this thread's copies) ...
        __syncthreads();                               // ... and everybody else's
        chunk_fma<SQ>(sw + (size_t)(kw0 + c * KC) * RR + 4 * rg, abuf + (c % NBUF) * (SBLK * KROW) + (sg * SQ) * KROW, warp, acc);
        __syncthreads();                               // buffer c % 3 is refilled by the issue of the next iteration
    }
    copy_wait_prior<0>();
}

// ---- tensor-core form of `segment` (rnn_fast: S = B*L sequences of n samples) ---------------------
// The fp32 FMA form above is bound by the shared-memory pipe (9 LDS.128 per 80 FMAs and lane); here the same chunk
// (128 k x <= 20 sequences, staged by the same cp.async ring) feeds mma.sync m16n8k16 with fp16 hi/lo split operands and
// fp32 accumulation (mma_frag.cuh, the scheme of every other GEMM of the library): the CTA's 32 gate rows are four
// n-tiles (one per gate), the sequences one or two m-tiles, a warp owns ONE 16-k step of every chunk and keeps its
// partial sums for the whole step (input half + recurrent half); per k-step 6 LDS.64 of activations + 4 LDS.128 of
// weight fragments for 24 MMAs.  Row stride 136 floats keeps the half-warps' 64-bit reads conflict-free.
constexpr int KROWT = KC + 8;
static_assert(NBUF * SBLK * KROWT * 4 <= 33 * 1024, "activation ring of the tensor-core form");
constexpr int ABUF_TC = NBUF * SBLK * KROWT;

// the MMAs of this warp's 16-k step of one staged chunk: buf = chunk + 16 warp + 2 t (row stride KROWT), wk = the k-step's fragments + lane
template <int MT>
__device__ __forceinline__ void chunk_mma(const uint4* __restrict__ wk, const float* __restrict__ buf, int g, int nseq,
                                          float (&big)[MT][4][4], float (&small)[MT][4][4]) {
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int r0 = 16 * m + g, r1 = r0 + 8;
        const float2 z = make_float2(0.f, 0.f);
        const float2 v0 = r0 < nseq ? *reinterpret_cast<const float2*>(buf + r0 * KROWT) : z;
        const float2 v1 = r1 < nseq ? *reinterpret_cast<const float2*>(buf + r1 * KROWT) : z;
        const float2 v2 = r0 < nseq ? *reinterpret_cast<const float2*>(buf + r0 * KROWT + 8) : z;
        const float2 v3 = r1 < nseq ? *reinterpret_cast<const float2*>(buf + r1 * KROWT + 8) : z;
        frag::split2(v0.x, v0.y, ah[m][0], al[m][0]);
        frag::split2(v1.x, v1.y, ah[m][1], al[m][1]);
        frag::split2(v2.x, v2.y, ah[m][2], al[m][2]);
        frag::split2(v3.x, v3.y, ah[m][3], al[m][3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 b = wk[j * 32];
#pragma unroll
        for (int m = 0; m < MT; ++m) frag::mma3(big[m][j], small[m][j], ah[m], al[m], b);
    }
}

// The recurrent half of the tensor-core form fed by tagged words (p.xchg): the chunk's nseq x 128 words go from L2 into
// registers (issued before the previous chunk's MMAs), are validated / re-read until they carry the step's tag, and land
// as plain floats in one of two chunk buffers.  One block barrier per chunk.
template <int MT>
__device__ __forceinline__ void recurrent_tc_tagged(const uint4* __restrict__ wf, float* __restrict__ abuf,
                                                    const unsigned long long* __restrict__ xsrc, int nseq, int ks0,
                                                    unsigned tag, unsigned* error, int tid, float (&big)[MT][4][4],
                                                    float (&small)[MT][4][4]) {
    constexpr int WPT = SBLK * KC / RT;               // words per thread and chunk at 20 sequences (10)
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    unsigned long long w[WPT];
    auto load = [&](int c) {
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
            const int i = tid + j * RT;
            if (i < nseq * KC) w[j] = ld_tagged(xsrc + (size_t)(i / KC) * kImuH + c * KC + (i % KC));
        }
    };
    auto commit = [&](int c, float* buf) {
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
            const int i = tid + j * RT;
            if (i < nseq * KC)
                buf[(i / KC) * KROWT + (i % KC)] = wait_tagged(xsrc + (size_t)(i / KC) * kImuH + c * KC + (i % KC), w[j], tag, error);
        }
    };
    load(0);
#pragma unroll 1
    for (int c = 0; c < kImuH / KC; ++c) {
        float* buf = abuf + (c & 1) * (SBLK * KROWT);
        commit(c, buf);
        __syncthreads();
        if (c + 1 < kImuH / KC) load(c + 1);
        chunk_mma<MT>(wf + (size_t)(ks0 + c * (KC / 16) + warp) * 4 * 32 + lane, buf + 16 * warp + 2 * t, g, nseq, big, small);
    }
    __syncthreads();                                   // the buffers double as the partial-sum area
}

// The same for the fp32 form at <= 4 sequences (rnn_slow): all 512 values of every sequence in ONE round trip
// (<= 8 words per thread), laid out as four chunks [c][4 rows][KROW].
__device__ __forceinline__ void recurrent_fp32_tagged(const float* __restrict__ sw, float* __restrict__ abuf,
                                                      const unsigned long long* __restrict__ xsrc, int nseq, int kw0,
                                                      unsigned tag, unsigned* error, int tid, float (&acc)[4][1]) {
    constexpr int WPT = 4 * kImuH / RT;               // 8
    const int lane = tid & 31, warp = tid >> 5, rg = lane >> 2, sg = lane & 3;
    unsigned long long w[WPT];
#pragma unroll
    for (int j = 0; j < WPT; ++j) {
        const int i = tid + j * RT;
        if (i < nseq * kImuH) w[j] = ld_tagged(xsrc + i);
    }
#pragma unroll
    for (int j = 0; j < WPT; ++j) {
        const int i = tid + j * RT;
        if (i < nseq * kImuH) {
            const int sq = i / kImuH, k = i % kImuH;
            abuf[((k / KC) * 4 + sq) * KROW + (k % KC)] = wait_tagged(xsrc + i, w[j], tag, error);
        }
    }
    __syncthreads();
    // (rows >= nseq of a chunk are never written: their lanes' sums are never read either)
#pragma unroll 1
    for (int c = 0; c < kImuH / KC; ++c)
        chunk_fma<1>(sw + (size_t)(kw0 + c * KC) * RR + 4 * rg, abuf + (c * 4 + sg) * KROW, warp, acc);
    __syncthreads();
}

template <int SQ>
struct TaggedFp32 {       // only the one-sequence-per-lane tile (S <= 4) has a tagged form
    static __device__ __forceinline__ void run(const float*, float*, const unsigned long long*, int, int, unsigned, unsigned*, int,
                                               float (&)[4][SQ]) {}
};
template <>
struct TaggedFp32<1> {
    static __device__ __forceinline__ void run(const float* sw, float* abuf, const unsigned long long* xsrc, int nseq, int kw0,
                                               unsigned tag, unsigned* error, int tid, float (&acc)[4][1]) {
        recurrent_fp32_tagged(sw, abuf, xsrc, nseq, kw0, tag, error, tid, acc);
    }
};

template <int MT>
__device__ __forceinline__ void segment_tc(const uint4* __restrict__ wf, float* __restrict__ abuf, const float* __restrict__ src,
                                           size_t stride, int nseq, int ks0, int len, int tid, float (&big)[MT][4][4],
                                           float (&small)[MT][4][4]) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int nchunk = len / KC;
    auto issue = [&](int c) {
        if (c < nchunk) {
            float* buf = abuf + (c % NBUF) * (SBLK * KROWT);
            for (int i = tid; i < nseq * (KC / 4); i += RT) {
                const int sq = i / (KC / 4), kq = i % (KC / 4);
                copy16_async(buf + sq * KROWT + 4 * kq, src + (size_t)sq * stride + c * KC + 4 * kq);
            }
        }
        copy_commit();
    };
    issue(0);
    issue(1);
    for (int c = 0; c < nchunk; ++c) {
        issue(c + 2);
        copy_wait_prior<2>();
        __syncthreads();
        chunk_mma<MT>(wf + (size_t)(ks0 + c * (KC / 16) + warp) * 4 * 32 + lane, abuf + (c % NBUF) * (SBLK * KROWT) + 16 * warp + 2 * t,
                      g, nseq, big, small);
        __syncthreads();
    }
    copy_wait_prior<0>();
}

// rnn_slow (one sequence per snippet: S = B, S * T <= kResMaxSeq): a step's input half x_t W_ih^T would read the
// CTA's whole W_ih slice (128 KB) from shared memory for ONE activation row per sequence, once per timestep and inside
// the lock-step cycle of the 64 CTAs (signal -> input half -> recurrent half -> cell).  The timesteps of a sequence are independent
// rows for that product, so it is done up front for all of them, 20 timesteps per pass through the 4 x 5 register tile
// that rnn_fast uses for 20 sequences: W_ih is read T/20 times instead of T times, and the per-step cycle keeps only the
// recurrent half.  gx [S][T][32] (this CTA's gate rows) lives in global scratch: written and read by this CTA only.
__device__ __forceinline__ void precompute_inputs(const ResParams& p, const float* sw, float* abuf, float* gx) {
    constexpr int SQ = 5, SP = 4 * SQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rg = lane >> 2, sg = lane & 3;
    for (int s = 0; s < p.S; ++s) {
        for (int t0 = 0; t0 < p.T; t0 += SP) {
            const int nrow = (p.T - t0) < SP ? (p.T - t0) : SP;
            float acc[4][SQ];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < SQ; ++j) acc[r][j] = 0.f;
            segment<SQ>(sw, abuf, p.x + ((size_t)s * p.T + t0) * p.In, (size_t)p.In, nrow, 0, p.In, tid, acc);
            float* red = abuf;                                   // [RW][SP][32]
#pragma unroll
            for (int j = 0; j < SQ; ++j)
                *reinterpret_cast<float4*>(red + ((size_t)(warp * SP + sg * SQ + j)) * RR + 4 * rg) =
                    make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            __syncthreads();
            for (int i = tid; i < nrow * RR; i += RT) {
                const int sl = i / RR, col = i % RR;
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < RW; ++w) v += red[(w * SP + sl) * RR + col];
                gx[((size_t)s * p.T + t0 + sl) * RR + col] = v;
            }
            __syncthreads();
        }
    }
}

template <int SQ>
__device__ __forceinline__ void run_steps(const ResParams& p, float* sw, float* abuf, int dir, int ug) {
    constexpr int SP = 4 * SQ;
    static_assert(SQ == 1 || SQ == 2 || SQ == 3 || SQ == 5, "sequence groups of 1, 2, 3 or 5");
    static_assert(RW * SP * RR <= ABUF, "partial-sum buffer");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rg = lane >> 2, sg = lane & 3;
    const int H2 = 2 * kImuH;
    const float* bsrc = p.bias + (dir * RGROUPS + ug) * RR;      // 32 floats, L1-resident
    float* gx = nullptr;
    if (p.gxs) {
        gx = p.gxs + (size_t)blockIdx.x * p.S * p.T * RR;
        if (p.t_begin == 0) precompute_inputs(p, sw, abuf, gx);  // (the emulator's later one-step launches find it in place)
    }
    for (int t = p.t_begin; t < p.t_end; ++t) {
        const int tt = dir ? p.T - 1 - t : t, tp = dir ? tt + 1 : tt - 1;
        const int nblocks = (p.S + SP - 1) / SP;
        bool waited = t == 0;
        for (int blk = 0; blk < nblocks; ++blk) {
            const int s0 = blk * SP, nseq = (p.S - s0) < SP ? (p.S - s0) : SP;
            float acc[4][SQ];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < SQ; ++j) acc[r][j] = 0.f;
            // ---- input half: x_t W_ih^T (independent of the other CTAs; runs before the wait for h_{t-1}) ------------
            if (!gx) segment<SQ>(sw, abuf, p.x + ((size_t)s0 * p.T + tt) * p.In, (size_t)p.T * p.In, nseq, 0, p.In, tid, acc);
            if (t > 0 && SQ == 1 && p.xchg) {
                // ---- recurrent half from tagged words: the load is the wait ----------------------------------------------
                TaggedFp32<SQ>::run(sw, abuf, p.xchg + ((size_t)(dir * 2 + ((t - 1) & 1)) * p.S + s0) * kImuH, nseq, p.In,
                                    (unsigned)t, p.error, tid, acc);
            } else if (t > 0) {
                if (!waited) {
                    // ---- wait for h_{t-1} of all 64 unit groups of this direction ------------------------------------
                    if (p.flags && tid == 0) {
                        volatile unsigned* f = p.flags + dir * p.T + (t - 1);
                        unsigned polls = 0;
                        while (*f < (unsigned)RGROUPS) {
#ifndef MMEGO_EMUL
                            __nanosleep(20);
#endif
                            if (++polls > (1u << 22)) {        // ~0.1 s: something is badly wrong -- do not hang the GPU
                                *p.error = 1u;
                                break;
                            }
                        }
                        __threadfence();
                    }
                    __syncthreads();
                    waited = true;
                }
                // ---- recurrent half: h_{t-1} W_hh^T (h is read through L2: other CTAs wrote it) ----------------------
                segment<SQ>(sw, abuf, p.y + ((size_t)s0 * p.T + tp) * H2 + dir * kImuH, (size_t)p.T * H2, nseq, p.In, kImuH,
                            tid, acc);
            }
            // ---- cross-warp sum of the K slices (the chunk buffers are free now), then the cell update ------------------
            float* red = abuf;                                   // [RW][SP][32]
#pragma unroll
            for (int j = 0; j < SQ; ++j)
                *reinterpret_cast<float4*>(red + ((size_t)(warp * SP + sg * SQ + j)) * RR + 4 * rg) =
                    make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            __syncthreads();
            for (int i = tid; i < nseq * RU; i += RT) {
                const int sl = i / RU, u = i % RU, s = s0 + sl;
                float pre[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float v = bsrc[g * RU + u];
                    if (gx) v += RES_LDCG(gx + ((size_t)s * p.T + tt) * RR + g * RU + u);
#pragma unroll
                    for (int w = 0; w < RW; ++w) v += red[(w * SP + sl) * RR + g * RU + u];
                    pre[g] = v;
                }
                const float gi = sigm(pre[0]), gf = sigm(pre[1]), gg = tanhf(pre[2]), go = sigm(pre[3]);
                float* cp = p.cstate + ((size_t)dir * p.S + s) * kImuH + ug * RU + u;
                const float cprev = t > 0 ? *cp : 0.f;
                const float cn = gf * cprev + gi * gg;
                *cp = cn;
                const float hn = go * tanhf(cn);
                p.y[((size_t)s * p.T + tt) * H2 + dir * kImuH + ug * RU + u] = hn;
                if (p.xchg) st_tagged(p.xchg + ((size_t)(dir * 2 + (t & 1)) * p.S + s) * kImuH + ug * RU + u, (unsigned)(t + 1), hn);
            }
            __syncthreads();                                     // red is rewritten by the next block / step
        }
        if (!p.xchg) {
            __threadfence();
            __syncthreads();
            if (p.flags && tid == 0) atomicAdd(p.flags + dir * p.T + t, 1u);
        }
    }
}

// Tensor-core form WITHOUT the shared-memory staging: once the product runs on mma.sync a chunk's MMAs take ~0.1 us, and
// the cp.async ring (two 10 KB chunks in flight, two block barriers per chunk) became the step: 12 chunks x ~0.7 us of L2
// latency for ~1.2 us of MMAs (ncu: tensor pipe 17-20 % active, profiles/r02g_resident_ncu_summary.txt).  Here a warp owns
// a CONTIGUOUS range of 16-k steps and reads its A fragments straight from global memory / L2 -- a lane's float2 is half
// of a 32-byte sector whose other half its neighbour lanes read, so no byte is fetched twice -- four k-steps of loads in
// flight per warp, no barrier anywhere in the K loop.  CG = the operand was written by other CTAs of this launch
// (h_{t-1}): read through L2 (ld.global.cg), after the arrival counter has been seen.
template <int MT, bool CG>
__device__ __forceinline__ void segment_tc_direct(const uint4* __restrict__ wf, const float* __restrict__ src, size_t stride,
                                                  int nseq, int ks0, int len, int tid, float (&big)[MT][4][4],
                                                  float (&small)[MT][4][4]) {
    constexpr int U = 4;                               // k-steps of loads in flight
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int per = len / 16 / RW;                     // k-steps per warp (len is a multiple of 128): 4 or 8
    const int kbase = warp * per;
    const float* rp[MT][2];
    bool live[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = 16 * m + 8 * h + g;
            live[m][h] = r < nseq;
            rp[m][h] = src + (size_t)(live[m][h] ? r : 0) * stride + 2 * t;
        }
    auto ld2 = [&](const float* ptr) {
#ifdef MMEGO_EMUL
        return *reinterpret_cast<const float2*>(ptr);
#else
        return CG ? __ldcg(reinterpret_cast<const float2*>(ptr)) : __ldg(reinterpret_cast<const float2*>(ptr));
#endif
    };
    for (int k0 = 0; k0 < per; k0 += U) {
        float2 v[U][MT][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int col = 16 * (kbase + k0 + u);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float2 z = make_float2(0.f, 0.f);
                v[u][m][0] = live[m][0] ? ld2(rp[m][0] + col) : z;
                v[u][m][1] = live[m][1] ? ld2(rp[m][1] + col) : z;
                v[u][m][2] = live[m][0] ? ld2(rp[m][0] + col + 8) : z;
                v[u][m][3] = live[m][1] ? ld2(rp[m][1] + col + 8) : z;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) frag::split2(v[u][m][i].x, v[u][m][i].y, ah[m][i], al[m][i]);
            const uint4* wk = wf + (size_t)(ks0 + kbase + k0 + u) * 4 * 32 + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 b = wk[j * 32];
#pragma unroll
                for (int m = 0; m < MT; ++m) frag::mma3(big[m][j], small[m][j], ah[m], al[m], b);
            }
        }
    }
}

// The timestep loop of the tensor-core form: same protocol as run_steps (input half before the wait for h_{t-1}, cell
// update by one thread per (sequence, unit), one arrival per CTA and step), blocks of up to 20 sequences.
template <int MT>
__device__ __forceinline__ void run_steps_tc(const ResParams& p, const uint4* wf, float* abuf, int dir, int ug) {
    constexpr int SP = SBLK;
    static_assert(RW * SP * RR <= ABUF_TC, "partial-sum buffer");
    static_assert(RW * 16 == KC, "one 16-k step of a chunk per warp");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int H2 = 2 * kImuH;
    const float* bsrc = p.bias + (dir * RGROUPS + ug) * RR;
    const float os = p.wscale[dir * RGROUPS + ug];
    for (int t = p.t_begin; t < p.t_end; ++t) {
        const int tt = dir ? p.T - 1 - t : t, tp = dir ? tt + 1 : tt - 1;
        const int nblocks = (p.S + SP - 1) / SP;
        bool waited = t == 0;
        for (int blk = 0; blk < nblocks; ++blk) {
            const int s0 = blk * SP, nseq = (p.S - s0) < SP ? (p.S - s0) : SP;
            float big[MT][4][4], small[MT][4][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) big[m][j][i] = small[m][j][i] = 0.f;
            if (p.direct)
                segment_tc_direct<MT, false>(wf, p.x + ((size_t)s0 * p.T + tt) * p.In, (size_t)p.T * p.In, nseq, 0, p.In, tid, big, small);
            else
                segment_tc<MT>(wf, abuf, p.x + ((size_t)s0 * p.T + tt) * p.In, (size_t)p.T * p.In, nseq, 0, p.In, tid, big, small);
            if (t > 0 && p.xchg) {
                recurrent_tc_tagged<MT>(wf, abuf, p.xchg + ((size_t)(dir * 2 + ((t - 1) & 1)) * p.S + s0) * kImuH, nseq, p.In / 16,
                                        (unsigned)t, p.error, tid, big, small);
            } else if (t > 0) {
                if (!waited) {
                    if (p.flags && tid == 0) {
                        volatile unsigned* f = p.flags + dir * p.T + (t - 1);
                        unsigned polls = 0;
                        while (*f < (unsigned)RGROUPS) {
#ifndef MMEGO_EMUL
                            __nanosleep(20);
#endif
                            if (++polls > (1u << 22)) {
                                *p.error = 1u;
                                break;
                            }
                        }
                        __threadfence();
                    }
                    __syncthreads();
                    waited = true;
                }
                if (p.direct)
                    segment_tc_direct<MT, true>(wf, p.y + ((size_t)s0 * p.T + tp) * H2 + dir * kImuH, (size_t)p.T * H2, nseq,
                                                p.In / 16, kImuH, tid, big, small);
                else
                    segment_tc<MT>(wf, abuf, p.y + ((size_t)s0 * p.T + tp) * H2 + dir * kImuH, (size_t)p.T * H2, nseq, p.In / 16,
                                   kImuH, tid, big, small);
            }
            // cross-warp sum of the k-steps: C fragment (rows g / g+8 of m-tile m, columns 2tq, 2tq+1 of gate j) -> [warp][sequence][32]
            float* red = abuf;
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r0 = 16 * m + g, r1 = r0 + 8;
                    if (r0 < nseq)
                        *reinterpret_cast<float2*>(red + ((size_t)(warp * SP + r0)) * RR + j * RU + 2 * tq) =
                            make_float2(big[m][j][0] + small[m][j][0], big[m][j][1] + small[m][j][1]);
                    if (r1 < nseq)
                        *reinterpret_cast<float2*>(red + ((size_t)(warp * SP + r1)) * RR + j * RU + 2 * tq) =
                            make_float2(big[m][j][2] + small[m][j][2], big[m][j][3] + small[m][j][3]);
                }
            __syncthreads();
            for (int i = tid; i < nseq * RU; i += RT) {
                const int sl = i / RU, u = i % RU, s = s0 + sl;
                float pre[4];
#pragma unroll
                for (int gt = 0; gt < 4; ++gt) {
                    float v = 0.f;
#pragma unroll
                    for (int w = 0; w < RW; ++w) v += red[(w * SP + sl) * RR + gt * RU + u];
                    pre[gt] = fmaf(v, os, bsrc[gt * RU + u]);
                }
                const float gi = sigm(pre[0]), gf = sigm(pre[1]), gg = tanhf(pre[2]), go = sigm(pre[3]);
                float* cp = p.cstate + ((size_t)dir * p.S + s) * kImuH + ug * RU + u;
                const float cprev = t > 0 ? *cp : 0.f;
                const float cn = gf * cprev + gi * gg;
                *cp = cn;
                const float hn = go * tanhf(cn);
                p.y[((size_t)s * p.T + tt) * H2 + dir * kImuH + ug * RU + u] = hn;
                if (p.xchg) st_tagged(p.xchg + ((size_t)(dir * 2 + (t & 1)) * p.S + s) * kImuH + ug * RU + u, (unsigned)(t + 1), hn);
            }
            __syncthreads();
        }
        if (!p.xchg) {
            __threadfence();
            __syncthreads();
            if (p.flags && tid == 0) atomicAdd(p.flags + dir * p.T + t, 1u);
        }
    }
}

__global__ void __launch_bounds__(RT, 1) lstm_resident_kernel(const ResParams p) {
    MMEGO_DYN_SMEM(float, sm);
    float* sw = sm;                                    // [K][32]
    float* abuf = sm + (size_t)p.K * RR;               // activation chunks [3][20][132], later [RW][SP][32] partial sums
    const int tid = threadIdx.x;
    const int dir = blockIdx.x / RGROUPS, ug = blockIdx.x % RGROUPS;
    {
        const float4* src = reinterpret_cast<const float4*>(p.w + ((size_t)(dir * RGROUPS + ug)) * p.K * RR);
        float4* dst = reinterpret_cast<float4*>(sw);
        for (int i = tid; i < p.K * RR / 4; i += RT) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (p.wscale) {        // tensor-core form: sw holds B fragments
        if (p.S <= 16) run_steps_tc<1>(p, reinterpret_cast<const uint4*>(sw), abuf, dir, ug);
        else run_steps_tc<2>(p, reinterpret_cast<const uint4*>(sw), abuf, dir, ug);
        return;
    }
    // register tile = 4 gate rows x SQ sequences per lane; a block of the K loop covers 4 SQ sequences
    if (p.S <= 4) run_steps<1>(p, sw, abuf, dir, ug);
    else if (p.S <= 8) run_steps<2>(p, sw, abuf, dir, ug);
    else if (p.S <= 12) run_steps<3>(p, sw, abuf, dir, ug);
    else run_steps<5>(p, sw, abuf, dir, ug);
}

// fc1 + ReLU in fp32 (Net/IMU_Net.py:79): imu [rows,15] -> u [rows,512]; w = [512][15] | bias [512]
__global__ void __launch_bounds__(256) res_fc1_kernel(const float* __restrict__ imu, const float* __restrict__ w,
                                                      float* __restrict__ u, long long rows) {
    __shared__ float sw[kImuH * kImuFeat + kImuH];
    for (int i = threadIdx.x; i < kImuH * kImuFeat + kImuH; i += 256) sw[i] = w[i];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < rows * kImuH; i += (long long)gridDim.x * 256) {
        const long long r = i / kImuH;
        const int c = (int)(i % kImuH);
        const float* x = imu + r * kImuFeat;
        float a = sw[kImuH * kImuFeat + c];
#pragma unroll
        for (int k = 0; k < kImuFeat; ++k) a = fmaf(sw[c * kImuFeat + k], x[k], a);
        u[i] = fmaxf(a, 0.f);
    }
}

}  // namespace

// [2 dirs][64 groups][K][32]: col = gate*8 + u  <->  torch row gate*512 + group*8 + u;  k < In: W_ih, else W_hh
void pack_resident_layer(const StateDict& sd, const std::string& prefix, int layer, int In, std::vector<float>& w,
                         std::vector<float>& bias) {
    const int H = kImuH, K = In + H;
    w.assign((size_t)2 * RGROUPS * K * RR, 0.f);
    bias.assign((size_t)2 * RGROUPS * RR, 0.f);
    const char* sfx[2] = {"", "_reverse"};
    for (int d = 0; d < 2; ++d) {
        const std::string k = "l" + std::to_string(layer) + sfx[d];
        const float* wih = sd.get(prefix + "weight_ih_" + k, (long long)4 * H * In);
        const float* whh = sd.get(prefix + "weight_hh_" + k, (long long)4 * H * H);
        const float* bih = sd.get(prefix + "bias_ih_" + k, 4 * H);
        const float* bhh = sd.get(prefix + "bias_hh_" + k, 4 * H);
        for (int ug = 0; ug < RGROUPS; ++ug) {
            float* dst = &w[((size_t)d * RGROUPS + ug) * K * RR];
            for (int col = 0; col < RR; ++col) {
                const int r = (col / RU) * H + ug * RU + col % RU;
                for (int kk = 0; kk < In; ++kk) dst[(size_t)kk * RR + col] = wih[(size_t)r * In + kk];
                for (int kk = 0; kk < H; ++kk) dst[(size_t)(In + kk) * RR + col] = whh[(size_t)r * H + kk];
                bias[((size_t)d * RGROUPS + ug) * RR + col] = bih[r] + bhh[r];
            }
        }
    }
}

// Tensor-core form of the same slices: per (direction, group) the [32 rows][K = In + 512] matrix [W_ih | W_hh] as mma.sync
// B fragments (n-tile = gate, k-steps in K order), ONE scale per slice (returned in `scale`), so both halves of a step
// share their accumulators.
void pack_resident_layer_tc(const StateDict& sd, const std::string& prefix, int layer, int In, std::vector<float>& w,
                            std::vector<float>& scale) {
    const int H = kImuH, K = In + H;
    w.assign((size_t)2 * RGROUPS * K * RR, 0.f);
    scale.assign((size_t)2 * RGROUPS, 1.f);
    const char* sfx[2] = {"", "_reverse"};
    std::vector<float> slice((size_t)RR * K);
    std::vector<int> kmap(K), nmap(RR);
    for (int k = 0; k < K; ++k) kmap[k] = k;
    for (int n = 0; n < RR; ++n) nmap[n] = n;
    for (int d = 0; d < 2; ++d) {
        const std::string k = "l" + std::to_string(layer) + sfx[d];
        const float* wih = sd.get(prefix + "weight_ih_" + k, (long long)4 * H * In);
        const float* whh = sd.get(prefix + "weight_hh_" + k, (long long)4 * H * H);
        for (int ug = 0; ug < RGROUPS; ++ug) {
            for (int col = 0; col < RR; ++col) {
                const int r = (col / RU) * H + ug * RU + col % RU;
                for (int kk = 0; kk < In; ++kk) slice[(size_t)col * K + kk] = wih[(size_t)r * In + kk];
                for (int kk = 0; kk < H; ++kk) slice[(size_t)col * K + In + kk] = whh[(size_t)r * H + kk];
            }
            scale[(size_t)d * RGROUPS + ug] =
                pack_mma_weight(slice.data(), K, kmap, nmap, &w[((size_t)d * RGROUPS + ug) * K * RR]);
        }
    }
}

size_t resident_xchg_words(int S) { return (size_t)2 * 2 * S * kImuH; }
size_t resident_gx_floats(int S, int T) { return (size_t)2 * RGROUPS * S * T * RR; }
size_t resident_smem_bytes(int K) { return ((size_t)K * RR + (ABUF_TC > ABUF ? ABUF_TC : ABUF)) * sizeof(float); }

#ifdef MMEGO_EMUL
bool resident_supported(int) { return true; }          // the emulator runs one timestep per launch: no co-residency needed
#else
bool resident_supported(int sm_count) { return sm_count >= 2 * RGROUPS; }   // all 128 CTAs must be co-resident (1 per SM)
#endif

// One bidirectional H=512 layer over T steps for S <= kResMaxSeq sequences, fp32: x [S][T][In] -> y [S][T][1024].
// w / wscale: the fp32 slices of pack_resident_layer and null, or the fragments and scales of pack_resident_layer_tc.
// xchg: resident_xchg_words(S) 64-bit words or null (then the fence + arrival-counter exchange is used).
// cstate: [2][S][512] floats, flags: [2][T] unsigned + 1 error word, gxs: resident_gx_floats(S, T) floats or null (all scratch).
// Which form runs never depends on S, so a snippet's result does not depend on its batch-mates or on how a caller chunks a batch.  Returns 0, or -1 on a launch error.
int launch_lstm_resident(const float* x, int In, float* y, const float* w, const float* wscale, const float* bias,
                         float* cstate, unsigned* flags, float* gxs, unsigned long long* xchg, int direct, int S, int T,
                         cudaStream_t st) {
    ResParams p{};
    p.direct = (wscale && direct) ? 1 : 0;
    p.x = x; p.y = y; p.w = w; p.wscale = wscale; p.bias = bias; p.cstate = cstate;
    p.gxs = wscale ? nullptr : gxs;      // up-front input projections: fp32 form only (the caller passes it for rnn_slow)
    // tagged-word exchange: the one-sequence-per-lane fp32 form (S <= 4: all of h in one round trip) and the tensor-core form
    // with a single block of sequences per step.  With several blocks the counter exchange is paid once per step while the
    // register-staged tagged chunks cost every block more than the cp.async ring (measured, IMU_Net per call: B = 1 0.582 ->
    // 0.530 ms, but B = 2 0.857 -> 0.913, B = 4 1.421 -> 1.688 with tagged words everywhere).
    // The direct tensor-core form reads h as plain floats after the arrival counter.
    p.xchg = ((wscale && S <= SBLK && !p.direct) || (!wscale && S <= 4)) ? xchg : nullptr;
    if (p.xchg && cudaMemsetAsync(p.xchg, 0, resident_xchg_words(S) * sizeof(unsigned long long), st) != cudaSuccess) return -1;
    p.S = S; p.T = T; p.In = In; p.K = In + kImuH;
    const size_t smem = resident_smem_bytes(p.K);
    static int attr_bytes[64] = {0};
    int d = 0;
    cudaGetDevice(&d);
    if (attr_bytes[d & 63] < (int)smem) {
        if (cudaFuncSetAttribute(lstm_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -1;
        attr_bytes[d & 63] = (int)smem;
    }
#ifdef MMEGO_EMUL
    // the emulator runs CTAs one after another: one launch per timestep, no inter-CTA waiting
    p.flags = nullptr;
    p.error = nullptr;
    for (int t = 0; t < T; ++t) {
        p.t_begin = t;
        p.t_end = t + 1;
        MMEGO_LAUNCH(lstm_resident_kernel, dim3(2 * RGROUPS), dim3(RT), smem, st, p);
    }
    return 0;
#else
    p.flags = flags;
    p.error = flags + 2 * T;
    p.t_begin = 0;
    p.t_end = T;
    if (cudaMemsetAsync(flags, 0, (size_t)(2 * T + 1) * sizeof(unsigned), st) != cudaSuccess) return -1;
    void* args[] = {const_cast<ResParams*>(&p)};
    ++t_launches;
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_resident_kernel), dim3(2 * RGROUPS), dim3(RT), args, smem,
                                       st) == cudaSuccess ? 0 : -1;
#endif
}

void launch_res_fc1(const float* imu, const float* w, float* u, long long rows, cudaStream_t st) {
    if (rows <= 0) return;
    long long blocks = (rows * kImuH + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    MMEGO_LAUNCH(res_fc1_kernel, dim3((unsigned)blocks), dim3(256), 0, st, imu, w, u, rows);
}

}  // namespace mmego
