// point_layout.h -- offsets (in floats) of the folded per-point MLP weight blobs staged in shared memory by
// point_upper.cu / lower_frame.cu and produced on the host by pack.cpp.
// Every layer is stored as W[Cout][CinPad] (row-major, CinPad = Cin rounded up to 4, zero padded) followed by b[Cout].
// BatchNorm (eval) is folded:  W' = s W,  b' = s (b - mean) + beta,  s = gamma / sqrt(var + 1e-5).
#pragma once

namespace mmego {

constexpr int pad4(int c) { return (c + 3) / 4 * 4; }

// ---- Upper: PointNet 6->8->16->24, cat x[0:4] -> 28, GlobalPointNet 28->32->48->64, attn 64->1
//      (Net/Upper_Net.py:242-301)
struct UpperPointLayout {
    static constexpr int C0 = 6, C1 = 8, C2 = 16, C3 = 24, C3C = 28, C4 = 32, C5 = 48, C6 = 64;
    static constexpr int W1 = 0;
    static constexpr int B1 = W1 + C1 * pad4(C0);
    static constexpr int W2 = B1 + C1;
    static constexpr int B2 = W2 + C2 * pad4(C1);
    static constexpr int W3 = B2 + C2;
    static constexpr int B3 = W3 + C3 * pad4(C2);
    static constexpr int W4 = B3 + C3;
    static constexpr int B4 = W4 + C4 * pad4(C3C);
    static constexpr int W5 = B4 + C4;
    static constexpr int B5 = W5 + C5 * pad4(C4);
    static constexpr int W6 = B5 + C5;
    static constexpr int B6 = W6 + C6 * pad4(C5);
    static constexpr int WA = B6 + C6;          // attn weight [64]
    static constexpr int BA = WA + C6;          // attn bias [1]
    static constexpr int TOTAL = (BA + 1 + 3) / 4 * 4;
};

// ---- Lower: BasePointNet 6->16->32->61 (Net/Lower_Net.py:40-72) + FusionModule to_q/to_k/to_v 64->64 (:83-85)
struct LowerFrameLayout {
    static constexpr int C0 = 6, C1 = 16, C2 = 32, C3 = 61, D = 64;
    static constexpr int W1 = 0;
    static constexpr int B1 = W1 + C1 * pad4(C0);
    static constexpr int W2 = B1 + C1;
    static constexpr int B2 = W2 + C2 * pad4(C1);
    static constexpr int W3 = B2 + C2;
    static constexpr int B3 = W3 + pad4(C3) * pad4(C2);   // rows padded to 64 (3 zero rows)
    static constexpr int WQ = B3 + pad4(C3);
    static constexpr int BQ = WQ + D * D;
    static constexpr int WK = BQ + D;
    static constexpr int BK = WK + D * D;
    static constexpr int WV = BK + D;
    static constexpr int BV = WV + D * D;
    static constexpr int TOTAL = BV + D;
};

// ------------------------------------------------------------------------------------------------
// Tensor-core (mma.sync m16n8k16, fp16 hi/lo split) packing of the same networks: offsets in 32-bit WORDS.
// A layer with KS k-steps and NT n-tiles is stored as KS*NT*32 uint4 {hi.b0, hi.b1, lo.b0, lo.b1} in fragment order
// (index (s*NT + j)*32 + lane, see mma_frag.cuh), scaled by a power of two 2^e (undone by OS[layer] = 2^-e), followed
// by its fp32 bias padded to NT*8.
constexpr int mma_frag_words(int ks, int nt) { return ks * nt * 32 * 4; }

struct UpperMmaLayout {
    // layer:            L1      L2      L3      L4 (in: feat24|x0..3|pad)  L5      L6
    static constexpr int KS1 = 1, NT1 = 1, KS2 = 1, NT2 = 2, KS3 = 1, NT3 = 3, KS4 = 2, NT4 = 4, KS5 = 2, NT5 = 6,
                         KS6 = 3, NT6 = 8;
    static constexpr int F1 = 0;
    static constexpr int BI1 = F1 + mma_frag_words(KS1, NT1);
    static constexpr int F2 = BI1 + NT1 * 8;
    static constexpr int BI2 = F2 + mma_frag_words(KS2, NT2);
    static constexpr int F3 = BI2 + NT2 * 8;
    static constexpr int BI3 = F3 + mma_frag_words(KS3, NT3);
    static constexpr int F4 = BI3 + NT3 * 8;
    static constexpr int BI4 = F4 + mma_frag_words(KS4, NT4);
    static constexpr int F5 = BI4 + NT4 * 8;
    static constexpr int BI5 = F5 + mma_frag_words(KS5, NT5);
    static constexpr int F6 = BI5 + NT5 * 8;
    static constexpr int BI6 = F6 + mma_frag_words(KS6, NT6);
    static constexpr int WA = BI6 + NT6 * 8;    // attn weight [64] fp32
    static constexpr int BA = WA + 64;          // attn bias
    static constexpr int OS = BA + 1;           // out scales [6]
    static constexpr int TOTAL = (OS + 6 + 3) / 4 * 4;
};

struct LowerMmaLayout {
    // BasePointNet 6->16->32->61(64), then to_q on P' = [feat61 | x y z], to_k / to_v on the 15(16) joint features
    static constexpr int KS1 = 1, NT1 = 2, KS2 = 1, NT2 = 4, KS3 = 2, NT3 = 8, KSP = 4, NTP = 8;
    static constexpr int F1 = 0;
    static constexpr int BI1 = F1 + mma_frag_words(KS1, NT1);
    static constexpr int F2 = BI1 + NT1 * 8;
    static constexpr int BI2 = F2 + mma_frag_words(KS2, NT2);
    static constexpr int F3 = BI2 + NT2 * 8;
    static constexpr int BI3 = F3 + mma_frag_words(KS3, NT3);
    static constexpr int FQ = BI3 + NT3 * 8;
    static constexpr int BQ = FQ + mma_frag_words(KSP, NTP);
    static constexpr int FK = BQ + 64;
    static constexpr int BK = FK + mma_frag_words(KSP, NTP);
    static constexpr int FV = BK + 64;
    static constexpr int BV = FV + mma_frag_words(KSP, NTP);
    static constexpr int OS = BV + 64;          // out scales [6]: L1, L2, L3, q, k, v
    static constexpr int TOTAL = (OS + 6 + 3) / 4 * 4;
};

}  // namespace mmego
