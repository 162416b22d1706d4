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

}  // namespace mmego
