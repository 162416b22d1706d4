// heads_mma.cu -- the small fully connected heads on mma.sync m16n8k16 fragments (mma_frag.cuh, fp16 hi/lo split):
//   Upper_Net MLPHead   128 -> 128 (ReLU) -> 87                              Net/Upper_Net.py:343-353
//   Lower_Net fusion    [rnn_pk out 128 | upper_h 45] -> 128 (ReLU) -> 64 (ReLU) -> 42     Net/Lower_Net.py:119-124
// One launch per head instead of one FFMA GEMM per layer: a warp owns 16 rows (frames) and carries them through all
// layers in registers; the weights of all layers live in shared memory as B fragments (110 / 135 KB, one CTA per SM).
// Layer 0 streams its input from global memory one k-step at a time (two passes over K, one per half of its 128
// outputs, so that only 8 accumulator tiles are live); the later layers read their A fragments from registers.
//
// The heads' outputs never reach HBM: the last layer's C fragments go to a per-warp shared-memory tile and the SAME
// kernel decodes them (TAIL 1: 14 x 6D -> rotations, forward kinematics, Transform2R -> upper_l; TAIL 2: the lower
// chain -> lower_l, plus the 21-joint assembly and the error sums of Processor/Test/Demo_test.py:121-123, 64-69 with a
// CTA-level float64 reduction and one atomic per CTA and accumulator).  These were four launch-latency-sized kernels
// (decode 36 + 13 us, assembly/metrics 56 us for ~60 MB: 0.1 of the HBM roofline) whose inputs and outputs
// round-tripped HBM; fused, their traffic is the compulsory one: hs/uh in, joints (and pred) out.
#include "internal.h"
#include "decode_math.cuh"
#include "mma_frag.cuh"
#include "point_layout.h"

namespace mmego {

namespace {

constexpr int HT = 256;     // threads per CTA (8 warps = 8 row tiles in flight)
constexpr int HW = HT / 32;

// per-warp staging row of the fused tails: [head output | joints | pred]; odd strides -> conflict-free per-frame access
constexpr int kUpStage = 87 + 45 + 1;            // 133
constexpr int kLoStage = 42 + 24 + 63;           // 129

// blob (32-bit words): layer l = frags [KS_l * NT_l * 32] uint4 | bias [NT_l * 8] fp32;  then out scales [3]
template <int KS0, int KS1, int NT1, int KS2, int NT2>
struct HeadLayout {
    static constexpr int NT0 = 16;
    static constexpr int F0 = 0;
    static constexpr int B0 = F0 + mma_frag_words(KS0, NT0);
    static constexpr int F1 = B0 + NT0 * 8;
    static constexpr int B1 = F1 + mma_frag_words(KS1, NT1);
    static constexpr int F2 = B1 + NT1 * 8;
    static constexpr int B2 = F2 + (NT2 ? mma_frag_words(KS2, NT2) : 0);
    static constexpr int OS = B2 + NT2 * 8;
    static constexpr int TOTAL = (OS + 3 + 3) / 4 * 4;
};

// C fragments -> fp32 rows of `out` (row stride ldo), columns < n_out
template <int NT>
__device__ __forceinline__ void store_rows(const float (&c)[NT][4], float* out, long long ldo, long long r0, long long r1,
                                           bool live0, bool live1, int tq, int n_out, int j0 = 0) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int col = 8 * (j0 + j) + 2 * tq;
        if (col < n_out) {
            if (live0) out[r0 * ldo + col] = c[j][0];
            if (live1) out[r1 * ldo + col] = c[j][2];
        }
        if (col + 1 < n_out) {
            if (live0) out[r0 * ldo + col + 1] = c[j][1];
            if (live1) out[r1 * ldo + col + 1] = c[j][3];
        }
    }
}

// x0 [F, K0A] (row stride ld0) and optionally x1 [F, K0B] (row stride ld1) form the input row [x0 | x1] (zero padded to
// 16 KS0).  THREE = third layer present.
// TAIL: 0 = plain head (rows of `out`), 1 = + Upper_Net decode, 2 = + Lower_Net decode (+ assembly + metrics)
template <int KS0, int KS1, int NT1, int KS2, int NT2, int TAIL>
__global__ void __launch_bounds__(HT, 1) head_mma_kernel(const float* __restrict__ x0, int ld0, int k0a,
                                                         const float* __restrict__ x1, int ld1, int k0b,
                                                         const float* __restrict__ blob, float* __restrict__ out,
                                                         int ldo, int n_out, long long F, const HeadTail tail) {
    using HL = HeadLayout<KS0, KS1, NT1, KS2, NT2>;
    constexpr int SLD = TAIL == 1 ? kUpStage : kLoStage;
    MMEGO_DYN_SMEM(uint32_t, sw);
    for (int i = threadIdx.x * 4; i < HL::TOTAL; i += HT * 4)
        *reinterpret_cast<uint4*>(sw + i) = *reinterpret_cast<const uint4*>(blob + i);
    float* stage = reinterpret_cast<float*>(sw + HL::TOTAL) + (threadIdx.x >> 5) * 16 * SLD;     // this warp's 16 rows
    double* acc = reinterpret_cast<double*>(sw + HL::TOTAL + HW * 16 * SLD);                      // [HW][46] (TAIL 2)
    if (TAIL == 2)
        for (int i = threadIdx.x; i < HW * dec::kSumsLen; i += HT) acc[i] = 0.0;
    __syncthreads();
    const uint4* wf = reinterpret_cast<const uint4*>(sw);
    const float* wfl = reinterpret_cast<const float*>(sw);
    const float os0 = wfl[HL::OS], os1 = wfl[HL::OS + 1], os2 = wfl[HL::OS + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
    auto in2 = [&](long long row, int c) -> float2 {        // input channels c, c+1 of `row`
        if (c + 1 < k0a) return *reinterpret_cast<const float2*>(x0 + row * ld0 + c);
        float2 v = make_float2(0.f, 0.f);
        if (c < k0a) v.x = x0[row * ld0 + c];
        else if (c - k0a < k0b) v.x = x1[row * ld1 + c - k0a];
        if (c + 1 >= k0a && c + 1 - k0a < k0b) v.y = x1[row * ld1 + c + 1 - k0a];
        return v;
    };
    for (long long t0 = ((long long)blockIdx.x * (HT / 32) + warp) * 16; t0 < F; t0 += (long long)gridDim.x * (HT / 32) * 16) {
        const long long r0 = t0 + g, r1 = t0 + g + 8;
        const bool live0 = r0 < F, live1 = r1 < F;
        // ---- layer 0: 16 KS0 -> 128, ReLU, in two halves of 8 n-tiles; its output becomes the A operand of layer 1 ----
        uint32_t a1h[KS1][4], a1l[KS1][4];
        static_assert(KS1 == 8, "layer 0 has 128 outputs");
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float big[8][4], small[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) big[j][i] = small[j][i] = 0.f;
#pragma unroll 2
            for (int s = 0; s < KS0; ++s) {
                uint32_t ah[4], al[4];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int c = 16 * s + 8 * hh + 2 * tq;
                    const float2 v0 = live0 ? in2(r0, c) : make_float2(0.f, 0.f);
                    const float2 v1 = live1 ? in2(r1, c) : make_float2(0.f, 0.f);
                    frag::split2(v0.x, v0.y, ah[2 * hh], al[2 * hh]);
                    frag::split2(v1.x, v1.y, ah[2 * hh + 1], al[2 * hh + 1]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) frag::mma3(big[j], small[j], ah, al, wf[HL::F0 / 4 + (s * 16 + 8 * half + j) * 32 + lane]);
            }
            float c0[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(wfl + HL::B0 + 8 * (8 * half + j) + 2 * tq);
#pragma unroll
                for (int i = 0; i < 4; ++i) c0[j][i] = fmaxf(fmaf(big[j][i] + small[j][i], os0, (i & 1) ? bv.y : bv.x), 0.f);
            }
            // n-tiles 8 half .. 8 half + 7 = k-steps 4 half .. 4 half + 3 of layer 1
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                frag::split2(c0[2 * s][0], c0[2 * s][1], a1h[4 * half + s][0], a1l[4 * half + s][0]);
                frag::split2(c0[2 * s][2], c0[2 * s][3], a1h[4 * half + s][1], a1l[4 * half + s][1]);
                frag::split2(c0[2 * s + 1][0], c0[2 * s + 1][1], a1h[4 * half + s][2], a1l[4 * half + s][2]);
                frag::split2(c0[2 * s + 1][2], c0[2 * s + 1][3], a1h[4 * half + s][3], a1l[4 * half + s][3]);
            }
        }
        if (NT2 == 0) {
            // ---- layer 1 is the last: no ReLU, straight to global (groups of <= 4 n-tiles keep the registers low) ----
#pragma unroll
            for (int j0 = 0; j0 < NT1; j0 += 4) {
                if (j0 + 4 <= NT1) {
                    float c1[4][4];
                    frag::dense_tile<KS1, 4, false, NT1>(wf + HL::F1 / 4, wfl + HL::B1, os1, a1h, a1l, c1, lane, j0);
                    if (TAIL) store_rows<4>(c1, stage, SLD, g, g + 8, true, true, tq, n_out, j0);
                    if (!TAIL || out) store_rows<4>(c1, out, ldo, r0, r1, live0, live1, tq, n_out, j0);
                } else {
                    constexpr int REM = NT1 % 4 ? NT1 % 4 : 4;
                    float c1[REM][4];
                    frag::dense_tile<KS1, REM, false, NT1>(wf + HL::F1 / 4, wfl + HL::B1, os1, a1h, a1l, c1, lane, j0);
                    if (TAIL) store_rows<REM>(c1, stage, SLD, g, g + 8, true, true, tq, n_out, j0);
                    if (!TAIL || out) store_rows<REM>(c1, out, ldo, r0, r1, live0, live1, tq, n_out, j0);
                }
            }
        } else {
            float c1[NT1][4];
            frag::dense_tile<KS1, NT1, true>(wf + HL::F1 / 4, wfl + HL::B1, os1, a1h, a1l, c1, lane);
            uint32_t a2h[KS2 ? KS2 : 1][4], a2l[KS2 ? KS2 : 1][4];
            frag::to_afrag<NT1, (KS2 ? KS2 : 1)>(c1, a2h, a2l);
            float c2[NT2 ? NT2 : 1][4];
            frag::dense_tile<(KS2 ? KS2 : 1), (NT2 ? NT2 : 1), false>(wf + HL::F2 / 4, wfl + HL::B2, os2, a2h, a2l, c2, lane);
            if (TAIL) store_rows<(NT2 ? NT2 : 1)>(c2, stage, SLD, g, g + 8, true, true, tq, n_out);
            if (!TAIL || out) store_rows<(NT2 ? NT2 : 1)>(c2, out, ldo, r0, r1, live0, live1, tq, n_out);
        }
        if (TAIL) {
            // ---- fused tail: lanes 0..15 decode one frame each out of the warp's staging tile ----
            __syncwarp();
            const long long row = t0 + lane;
            const bool mine = lane < 16 && row < F;
            float* sr = stage + (lane & 15) * SLD;
            constexpr int NIN = TAIL == 1 ? 87 : 42, NJ = TAIL == 1 ? 45 : 24;
            float vals[dec::kSumsLen];
            if (TAIL == 2) {
#pragma unroll
                for (int i = 0; i < dec::kSumsLen; ++i) vals[i] = 0.f;
            }
            if (mine) {
                const long long r = tail.row_offset + row;
                const long long bi = (tail.mode == 0) ? (r % tail.B_global) : (r / tail.L);
                const float* bd = tail.body + bi * 60;
                float rt[12];
#pragma unroll
                for (int k = 0; k < 9; ++k) rt[k] = tail.R[row * 9 + k];
#pragma unroll
                for (int k = 0; k < 3; ++k) rt[9 + k] = tail.t[row * 3 + k];
                float* q = tail.q ? tail.q + row * (TAIL == 1 ? 126 : 54) : nullptr;
                if (TAIL == 1) dec::upper_frame_decode(sr, bd, rt, sr + NIN, q);
                else dec::lower_frame_decode(sr, bd, rt, sr + NIN, q);
                if (TAIL == 2 && tail.assemble) {
                    const float* u = tail.upper_l + row * 45;
                    dec::assemble_frame(u, sr + NIN, sr + NIN + NJ);
                    if (tail.target && tail.sums) dec::frame_metrics(sr + NIN + NJ, tail.target + row * 63, u, sr + NIN, vals);
                }
            }
            __syncwarp();
            // coalesced write-back: the 16 rows of joints (and of pred) are contiguous in global memory
            const int nrow = (F - t0) < 16 ? (int)(F - t0) : 16;
            for (int i = lane; i < nrow * NJ; i += 32) tail.l[t0 * NJ + i] = stage[(i / NJ) * SLD + NIN + i % NJ];
            if (TAIL == 2 && tail.assemble && tail.pred)
                for (int i = lane; i < nrow * 63; i += 32) tail.pred[t0 * 63 + i] = stage[(i / 63) * SLD + NIN + NJ + i % 63];
            if (TAIL == 2 && tail.assemble && tail.target && tail.sums) {
                double* aw = acc + warp * dec::kSumsLen;
#pragma unroll
                for (int i = 0; i < dec::kSumsLen; ++i) {
                    float v = vals[i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) aw[i] += (double)v;
                }
            }
            __syncwarp();
        }
    }
    if (TAIL == 2) {
        // one float64 atomic per CTA and accumulator
        __syncthreads();
        if (tail.assemble && tail.target && tail.sums && threadIdx.x < dec::kSumsLen) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < HW; ++w) v += acc[w * dec::kSumsLen + threadIdx.x];
            if (v != 0.0) atomicAdd(&tail.sums[threadIdx.x], v);
        }
    }
}

template <int KS0, int KS1, int NT1, int KS2, int NT2, int TAIL>
void launch_head(const float* x0, int ld0, int k0a, const float* x1, int ld1, int k0b, const float* blob, float* out, int ldo,
                 int n_out, long long F, const HeadTail& tail, int sm_count, cudaStream_t st) {
    using HL = HeadLayout<KS0, KS1, NT1, KS2, NT2>;
    size_t smem = (size_t)HL::TOTAL * 4;
    if (TAIL) smem += (size_t)HW * 16 * (TAIL == 1 ? kUpStage : kLoStage) * 4 + 8 /*align*/ + (size_t)HW * dec::kSumsLen * 8;
    static bool attr_set[64] = {false};
    if (first_use_on_device(attr_set))
        cudaFuncSetAttribute(head_mma_kernel<KS0, KS1, NT1, KS2, NT2, TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long tiles = (F + 127) / 128;
    const long long grid = tiles < sm_count ? tiles : sm_count;
    MMEGO_LAUNCH((head_mma_kernel<KS0, KS1, NT1, KS2, NT2, TAIL>), dim3((unsigned)grid), dim3(HT), smem, st, x0, ld0, k0a, x1,
                 ld1, k0b, blob, out, ldo, n_out, F, tail);
}

}  // namespace

// Upper_Net MLPHead + decode: x [F,128] -> upper_l [F,15,3] (+ q [F,14,3,3]); o [F,87] is written only when o != null
void launch_upper_tail_mma(const float* x, const float* blob, float* o, long long F, const HeadTail& tail, int sm_count,
                           cudaStream_t st) {
    if (F <= 0) return;
    launch_head<8, 8, 11, 0, 0, 1>(x, 128, 128, nullptr, 0, 0, blob, o, 87, 87, F, tail, sm_count, st);
}
// Lower_Net fusion.fc0/fc1/fc2 + decode (+ assembly + metrics): [hs [F,128] | uh [F,45]] -> lower_l [F,8,3] (+ q, pred, sums)
void launch_lower_tail_mma(const float* hs, const float* uh, const float* blob, float* o, long long F, const HeadTail& tail,
                           int sm_count, cudaStream_t st) {
    if (F <= 0) return;
    launch_head<11, 8, 8, 4, 6, 2>(hs, 128, 128, uh, 45, 45, blob, o, 42, 42, F, tail, sm_count, st);
}

size_t upper_head_mma_words() { return HeadLayout<8, 8, 11, 0, 0>::TOTAL; }
size_t lower_head_mma_words() { return HeadLayout<11, 8, 8, 4, 6>::TOTAL; }

}  // namespace mmego
