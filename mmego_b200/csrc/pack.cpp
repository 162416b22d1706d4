// pack.cpp -- host-side weight packing: BatchNorm folding, K-segment padding, LSTM gate interleaving.
// Layout contracts are documented next to the kernels that consume them (gemm_ffma.cu, lstm_small.cu, point_layout.h).
#include "pack.h"

#include <cmath>
#include <cstring>

#include "point_layout.h"

namespace mmego {

BnAffine bn_affine(const StateDict& sd, const std::string& p, int C) {
    const float* g = sd.get(p + ".weight", C);
    const float* b = sd.get(p + ".bias", C);
    const float* m = sd.get(p + ".running_mean", C);
    const float* v = sd.get(p + ".running_var", C);
    BnAffine r;
    r.s.resize(C);
    r.o.resize(C);
    for (int c = 0; c < C; ++c) {
        r.s[c] = g[c] / std::sqrt(v[c] + 1e-5f);
        r.o[c] = b[c] - m[c] * r.s[c];
    }
    return r;
}

HostPackedGemm pack_linear(const float* W, const float* b, int N, const std::vector<int>& ksegs, const float* row_scale,
                           const float* row_offset) {
    HostPackedGemm g;
    g.N = N;
    g.nseg = (int)ksegs.size();
    int ktot = 0, kp = 0;
    for (int s = 0; s < g.nseg; ++s) {
        g.k[s] = ksegs[s];
        g.kpad[s] = pad16(ksegs[s]);
        ktot += ksegs[s];
        kp += g.kpad[s];
    }
    g.ldw = kp;
    g.w.assign((size_t)N * kp, 0.f);
    g.bias.assign(N, 0.f);
    for (int n = 0; n < N; ++n) {
        const float sc = row_scale ? row_scale[n] : 1.f;
        int src = 0, dst = 0;
        for (int s = 0; s < g.nseg; ++s) {
            for (int k = 0; k < g.k[s]; ++k) g.w[(size_t)n * kp + dst + k] = sc * W[(size_t)n * ktot + src + k];
            src += g.k[s];
            dst += g.kpad[s];
        }
        g.bias[n] = sc * (b ? b[n] : 0.f) + (row_offset ? row_offset[n] : 0.f);
    }
    return g;
}

// H=512 LSTM layer, one GEMM per direction: rows gate-interleaved per tile of 32 hidden units
//   packed row p = tile*128 + gate*32 + e   <->   torch row gate*H + tile*32 + e      (gates i,f,g,o)
// K segments: [input (In) | recurrent (H)].
HostBigLstm pack_big_lstm(const StateDict& sd, const std::string& prefix, int layer, int In, int H) {
    HostBigLstm out;
    const char* sfx[2] = {"", "_reverse"};
    for (int d = 0; d < 2; ++d) {
        const std::string k = "l" + std::to_string(layer) + sfx[d];
        const float* wih = sd.get(prefix + "weight_ih_" + k, (long long)4 * H * In);
        const float* whh = sd.get(prefix + "weight_hh_" + k, (long long)4 * H * H);
        const float* bih = sd.get(prefix + "bias_ih_" + k, 4 * H);
        const float* bhh = sd.get(prefix + "bias_hh_" + k, 4 * H);
        HostPackedGemm& g = out.dir[d];
        g.N = 4 * H;
        g.nseg = 2;
        g.k[0] = In; g.kpad[0] = pad16(In);
        g.k[1] = H;  g.kpad[1] = pad16(H);
        g.ldw = g.kpad[0] + g.kpad[1];
        g.w.assign((size_t)g.N * g.ldw, 0.f);
        g.bias.assign(g.N, 0.f);
        const int TU = 32;
        for (int p = 0; p < 4 * H; ++p) {
            const int tile = p / (4 * TU), gate = (p % (4 * TU)) / TU, e = p % TU;
            const int r = gate * H + tile * TU + e;
            std::memcpy(&g.w[(size_t)p * g.ldw], wih + (size_t)r * In, sizeof(float) * In);
            std::memcpy(&g.w[(size_t)p * g.ldw + g.kpad[0]], whh + (size_t)r * H, sizeof(float) * H);
            g.bias[p] = bih[r] + bhh[r];
        }
    }
    return out;
}

// H=64 LSTM layer: recurrent-kernel thread order tid = w*32 + gate*8 + e  <->  torch row gate*64 + w*8 + e
HostSmallLstm pack_small_lstm(const StateDict& sd, const std::string& prefix, int layer, int In) {
    const int H = kSmallH;
    HostSmallLstm out;
    out.in = In;
    HostPackedGemm& g = out.ih;
    g.N = 2 * 4 * H;
    g.nseg = 1;
    g.k[0] = In; g.kpad[0] = pad16(In);
    g.ldw = g.kpad[0];
    g.w.assign((size_t)g.N * g.ldw, 0.f);
    g.bias.assign(g.N, 0.f);
    out.whh.assign((size_t)2 * 4 * H * H, 0.f);
    const char* sfx[2] = {"", "_reverse"};
    for (int d = 0; d < 2; ++d) {
        const std::string k = "l" + std::to_string(layer) + sfx[d];
        const float* wih = sd.get(prefix + "weight_ih_" + k, (long long)4 * H * In);
        const float* whh = sd.get(prefix + "weight_hh_" + k, (long long)4 * H * H);
        const float* bih = sd.get(prefix + "bias_ih_" + k, 4 * H);
        const float* bhh = sd.get(prefix + "bias_hh_" + k, 4 * H);
        for (int tid = 0; tid < 4 * H; ++tid) {
            const int w = tid / 32, gate = (tid % 32) / 8, e = tid % 8;
            const int r = gate * H + w * 8 + e;
            const int n = d * 4 * H + tid;
            std::memcpy(&g.w[(size_t)n * g.ldw], wih + (size_t)r * In, sizeof(float) * In);
            g.bias[n] = bih[r] + bhh[r];
            std::memcpy(&out.whh[((size_t)d * 4 * H + tid) * H], whh + (size_t)r * H, sizeof(float) * H);
        }
    }
    return out;
}

namespace {
// conv1d(k=1) [Cout][Cin][1] + BN folded into W[rows][CinPad] + b[rows]
void fold_conv_bn(const StateDict& sd, const std::string& conv, const std::string& bn, int Cin, int Cout, float* W,
                  float* b, int cin_pad) {
    const float* w = sd.get(conv + ".weight", (long long)Cout * Cin);
    const float* bb = sd.get(conv + ".bias", Cout);
    BnAffine a = bn_affine(sd, bn, Cout);
    for (int o = 0; o < Cout; ++o) {
        for (int c = 0; c < Cin; ++c) W[o * cin_pad + c] = a.s[o] * w[o * Cin + c];
        b[o] = a.s[o] * bb[o] + a.o[o];
    }
}
}  // namespace

std::vector<float> pack_upper_point(const StateDict& sd) {
    using UL = UpperPointLayout;
    std::vector<float> v(UL::TOTAL, 0.f);
    fold_conv_bn(sd, "module0.conv1", "module0.cb1", 6, 8, &v[UL::W1], &v[UL::B1], pad4(6));
    fold_conv_bn(sd, "module0.conv2", "module0.cb2", 8, 16, &v[UL::W2], &v[UL::B2], pad4(8));
    fold_conv_bn(sd, "module0.conv3", "module0.cb3", 16, 24, &v[UL::W3], &v[UL::B3], pad4(16));
    fold_conv_bn(sd, "module1.gpointnet.conv1", "module1.gpointnet.cb1", 28, 32, &v[UL::W4], &v[UL::B4], pad4(28));
    fold_conv_bn(sd, "module1.gpointnet.conv2", "module1.gpointnet.cb2", 32, 48, &v[UL::W5], &v[UL::B5], pad4(32));
    fold_conv_bn(sd, "module1.gpointnet.conv3", "module1.gpointnet.cb3", 48, 64, &v[UL::W6], &v[UL::B6], pad4(48));
    std::memcpy(&v[UL::WA], sd.get("module1.gpointnet.attn.weight", 64), sizeof(float) * 64);
    v[UL::BA] = sd.get("module1.gpointnet.attn.bias", 1)[0];
    return v;
}

std::vector<float> pack_lower_frame(const StateDict& sd) {
    using LL = LowerFrameLayout;
    std::vector<float> v(LL::TOTAL, 0.f);
    const std::string p = "pointEncoder.module0.";
    fold_conv_bn(sd, p + "conv1", p + "cb1", 6, 16, &v[LL::W1], &v[LL::B1], pad4(6));
    fold_conv_bn(sd, p + "conv2", p + "cb2", 16, 32, &v[LL::W2], &v[LL::B2], pad4(16));
    fold_conv_bn(sd, p + "conv3", p + "cb3", 32, 61, &v[LL::W3], &v[LL::B3], pad4(32));
    const float* wq = sd.get("fusion.to_q.weight", 64 * 64);
    std::memcpy(&v[LL::WQ], wq, sizeof(float) * 64 * 64);
    std::memcpy(&v[LL::BQ], sd.get("fusion.to_q.bias", 64), sizeof(float) * 64);
    const float* wk = sd.get("fusion.to_k.weight", 64 * 64);
    const float* wv = sd.get("fusion.to_v.weight", 64 * 64);
    for (int o = 0; o < 64; ++o)
        for (int c = 0; c < 64; ++c) {   // transposed [c][o] for conflict-free per-output-channel reads
            v[LL::WK + c * 64 + o] = wk[o * 64 + c];
            v[LL::WV + c * 64 + o] = wv[o * 64 + c];
        }
    std::memcpy(&v[LL::BK], sd.get("fusion.to_k.bias", 64), sizeof(float) * 64);
    std::memcpy(&v[LL::BV], sd.get("fusion.to_v.bias", 64), sizeof(float) * 64);
    return v;
}

// ------------------------------------------------------------------------------------------------ mma.sync packing
namespace {
// IEEE binary16 <-> binary32 on the host (round to nearest even, saturating to +-65504), bit-exact with the device's
// cvt.rn.satfinite.f16.f32
uint16_t f32_to_f16_bits(float x) {
    if (!(x == x)) return 0x7e00;
    if (x > 65504.f) x = 65504.f;
    if (x < -65504.f) x = -65504.f;
    uint32_t u;
    std::memcpy(&u, &x, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    u &= 0x7fffffffu;
    if (u >= 0x38800000u) {                       // normal half: rebias the exponent, round 13 dropped bits
        uint32_t r = u - 0x38000000u;
        const uint32_t rem = r & 0x1fffu;
        r >>= 13;
        if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
        return (uint16_t)(sign | r);
    }
    if (u < 0x33000000u) return (uint16_t)sign;   // < 2^-25: rounds to zero
    const int e = (int)(u >> 23);                 // subnormal half
    const uint32_t m = (u & 0x7fffffu) | 0x800000u;
    const int shift = 126 - e;                    // 14..24
    uint32_t r = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (r & 1u))) ++r;
    return (uint16_t)(sign | r);
}
float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    float f;
    if (e == 0) {
        f = std::ldexp((float)m, -24);
    } else if (e == 31) {
        f = m ? NAN : INFINITY;
    } else {
        f = std::ldexp((float)(m | 0x400u), (int)e - 25);
    }
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u |= sign;
    std::memcpy(&f, &u, 4);
    return f;
}
float word_as_float(uint32_t w) {
    float f;
    std::memcpy(&f, &w, 4);
    return f;
}
}  // namespace

// B operand of  out[n] = sum_k in[k] * W[nmap[n]][kmap[k]]  (entries mapped to -1 are zero), see point_layout.h
float pack_mma_weight(const float* W, int ldw, const std::vector<int>& kmap, const std::vector<int>& nmap, float* out) {
    const int KS = (int)kmap.size() / 16, NT = (int)nmap.size() / 8;
    auto V = [&](int k, int n) { return (kmap[k] >= 0 && nmap[n] >= 0) ? W[(size_t)nmap[n] * ldw + kmap[k]] : 0.f; };
    float mx = 0.f;
    for (int k = 0; k < KS * 16; ++k)
        for (int n = 0; n < NT * 8; ++n) mx = std::max(mx, std::fabs(V(k, n)));
    int e = 0;
    if (mx > 0.f) {
        int ex;
        std::frexp(mx, &ex);            // mx = f * 2^ex, f in [0.5, 1)
        e = 14 - ex;                    // scaled max in [2^13, 2^14)
        e = std::max(-24, std::min(e, 40));
    }
    for (int s = 0; s < KS; ++s)
        for (int j = 0; j < NT; ++j)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3, n = 8 * j + g;
                uint32_t w[4];
                for (int r = 0; r < 2; ++r) {
                    uint16_t hi[2], lo[2];
                    for (int i = 0; i < 2; ++i) {
                        const float v = std::ldexp(V(16 * s + 2 * t + 8 * r + i, n), e);
                        hi[i] = f32_to_f16_bits(v);
                        lo[i] = f32_to_f16_bits(v - f16_bits_to_f32(hi[i]));
                    }
                    w[r] = (uint32_t)hi[0] | ((uint32_t)hi[1] << 16);
                    w[2 + r] = (uint32_t)lo[0] | ((uint32_t)lo[1] << 16);
                }
                float* dst = out + ((size_t)(s * NT + j) * 32 + lane) * 4;
                for (int i = 0; i < 4; ++i) dst[i] = word_as_float(w[i]);
            }
    return std::ldexp(1.0f, -e);
}

namespace {
std::vector<int> iota_map(int n, int pad_to, int first = 0) {
    std::vector<int> m(pad_to, -1);
    for (int i = 0; i < n; ++i) m[i] = first + i;
    return m;
}
}  // namespace

std::vector<float> pack_upper_point_mma(const std::vector<float>& f) {
    using UL = UpperPointLayout;
    using UM = UpperMmaLayout;
    std::vector<float> v(UM::TOTAL, 0.f);
    struct Lyr { int W, B, cin_pad, cout, F, BI, KS, NT; std::vector<int> kmap; };
    std::vector<int> k4(32, -1);                 // layer 4 input order: feat 0..23 | x[0:4] | 4 pad  (source order: x[0:4] | feat)
    for (int i = 0; i < 24; ++i) k4[i] = 4 + i;
    for (int i = 0; i < 4; ++i) k4[24 + i] = i;
    const Lyr L[6] = {
        {UL::W1, UL::B1, pad4(UL::C0), UL::C1, UM::F1, UM::BI1, UM::KS1, UM::NT1, iota_map(UL::C0, 16)},
        {UL::W2, UL::B2, pad4(UL::C1), UL::C2, UM::F2, UM::BI2, UM::KS2, UM::NT2, iota_map(UL::C1, 16)},
        {UL::W3, UL::B3, pad4(UL::C2), UL::C3, UM::F3, UM::BI3, UM::KS3, UM::NT3, iota_map(UL::C2, 16)},
        {UL::W4, UL::B4, pad4(UL::C3C), UL::C4, UM::F4, UM::BI4, UM::KS4, UM::NT4, k4},
        {UL::W5, UL::B5, pad4(UL::C4), UL::C5, UM::F5, UM::BI5, UM::KS5, UM::NT5, iota_map(UL::C4, 32)},
        {UL::W6, UL::B6, pad4(UL::C5), UL::C6, UM::F6, UM::BI6, UM::KS6, UM::NT6, iota_map(UL::C5, 48)},
    };
    for (int l = 0; l < 6; ++l) {
        v[UM::OS + l] = pack_mma_weight(&f[L[l].W], L[l].cin_pad, L[l].kmap, iota_map(L[l].cout, L[l].NT * 8), &v[L[l].F]);
        for (int o = 0; o < L[l].cout; ++o) v[L[l].BI + o] = f[L[l].B + o];
    }
    for (int c = 0; c < 64; ++c) v[UM::WA + c] = f[UL::WA + c];
    v[UM::BA] = f[UL::BA];
    return v;
}

std::vector<float> pack_lower_frame_mma(const std::vector<float>& f) {
    using LL = LowerFrameLayout;
    using LM = LowerMmaLayout;
    std::vector<float> v(LM::TOTAL, 0.f);
    v[LM::OS + 0] = pack_mma_weight(&f[LL::W1], pad4(LL::C0), iota_map(LL::C0, 16), iota_map(LL::C1, 16), &v[LM::F1]);
    v[LM::OS + 1] = pack_mma_weight(&f[LL::W2], pad4(LL::C1), iota_map(LL::C1, 16), iota_map(LL::C2, 32), &v[LM::F2]);
    v[LM::OS + 2] = pack_mma_weight(&f[LL::W3], pad4(LL::C2), iota_map(LL::C2, 32), iota_map(LL::C3, 64), &v[LM::F3]);
    for (int o = 0; o < LL::C1; ++o) v[LM::BI1 + o] = f[LL::B1 + o];
    for (int o = 0; o < LL::C2; ++o) v[LM::BI2 + o] = f[LL::B2 + o];
    for (int o = 0; o < LL::C3; ++o) v[LM::BI3 + o] = f[LL::B3 + o];
    // to_q reads P' = [feat 0..60 | x y z]; the reference's P is [x y z | feat] (Net/Lower_Net.py:70-71)
    std::vector<int> kq(64);
    for (int i = 0; i < 61; ++i) kq[i] = 3 + i;
    for (int i = 0; i < 3; ++i) kq[61 + i] = i;
    v[LM::OS + 3] = pack_mma_weight(&f[LL::WQ], 64, kq, iota_map(64, 64), &v[LM::FQ]);
    // the fp32 blob stores to_k / to_v transposed ([c][o]); undo that here
    std::vector<float> wk(64 * 64), wv(64 * 64);
    for (int o = 0; o < 64; ++o)
        for (int c = 0; c < 64; ++c) {
            wk[o * 64 + c] = f[LL::WK + c * 64 + o];
            wv[o * 64 + c] = f[LL::WV + c * 64 + o];
        }
    v[LM::OS + 4] = pack_mma_weight(wk.data(), 64, iota_map(64, 64), iota_map(64, 64), &v[LM::FK]);
    v[LM::OS + 5] = pack_mma_weight(wv.data(), 64, iota_map(64, 64), iota_map(64, 64), &v[LM::FV]);
    for (int o = 0; o < 64; ++o) {
        v[LM::BQ + o] = f[LL::BQ + o];
        v[LM::BK + o] = f[LL::BK + o];
        v[LM::BV + o] = f[LL::BV + o];
    }
    return v;
}

// H=64 bi-LSTM layer for lstm_small.cu's mma.sync kernels.  Gate column order (both GEMMs): n-tile j = 4u + gate holds
// units 8u..8u+7 of gate (i, f, g, o), so the four gates of a cell sit in the same lane of four neighbouring n-tiles.
//   blob = ih frags [2 dirs][KS=In/16][32 n-tiles][32 lanes] uint4 | bias [2][256] (b_ih + b_hh) | hh frags [2][4][32][32]
//          uint4 | out scales {ih d0, ih d1, hh d0, hh d1}
std::vector<float> pack_small_lstm_mma(const StateDict& sd, const std::string& prefix, int layer, int In) {
    const int H = kSmallH, KS = In / 16;
    if (In % 16) throw std::runtime_error("pack_small_lstm_mma: In must be a multiple of 16");
    const size_t ihw = (size_t)mma_frag_words(KS, 32), hhw = (size_t)mma_frag_words(4, 32);
    std::vector<float> v(2 * ihw + 512 + 2 * hhw + 4, 0.f);
    std::vector<int> nmap(256);
    for (int j = 0; j < 32; ++j)
        for (int n = 0; n < 8; ++n) nmap[8 * j + n] = (j & 3) * H + 8 * (j >> 2) + n;
    const char* sfx[2] = {"", "_reverse"};
    for (int d = 0; d < 2; ++d) {
        const std::string k = "l" + std::to_string(layer) + sfx[d];
        const float* wih = sd.get(prefix + "weight_ih_" + k, (long long)4 * H * In);
        const float* whh = sd.get(prefix + "weight_hh_" + k, (long long)4 * H * H);
        const float* bih = sd.get(prefix + "bias_ih_" + k, 4 * H);
        const float* bhh = sd.get(prefix + "bias_hh_" + k, 4 * H);
        float* tail = &v[2 * ihw + 512 + 2 * hhw];
        tail[d] = pack_mma_weight(wih, In, iota_map(In, In), nmap, &v[d * ihw]);
        tail[2 + d] = pack_mma_weight(whh, H, iota_map(H, H), nmap, &v[2 * ihw + 512 + d * hhw]);
        for (int c = 0; c < 256; ++c) v[2 * ihw + d * 256 + c] = bih[nmap[c]] + bhh[nmap[c]];
    }
    return v;
}

// IMU_Net fc1 (15 -> 512) for imu_fc1_mma_kernel: one k-step, 64 n-tiles.  Output channels are permuted inside every
// group of four n-tiles so that a lane's eight accumulators of a row are eight CONSECUTIVE channels (one 16-byte store):
//   column n of n-tile j  <->  channel 32 (j/4) + 8 (n/2) + 2 (j%4) + n%2
//   blob = frags [64][32] uint4 | bias [512] in column order | out scale
std::vector<float> pack_imu_fc1_mma(const HostPackedGemm& g) {
    const int H = kImuH;
    std::vector<float> v((size_t)mma_frag_words(1, 64) + H + 4, 0.f);
    std::vector<int> nmap(H);
    for (int j = 0; j < 64; ++j)
        for (int n = 0; n < 8; ++n) nmap[8 * j + n] = 32 * (j >> 2) + 8 * (n >> 1) + 2 * (j & 3) + (n & 1);
    v[mma_frag_words(1, 64) + H] = pack_mma_weight(g.w.data(), g.ldw, iota_map(kImuFeat, 16), nmap, v.data());
    for (int c = 0; c < H; ++c) v[mma_frag_words(1, 64) + c] = g.bias[nmap[c]];
    return v;
}

// Fully connected heads for heads_mma.cu: layers given as (W [N][K] row-major, b [N]); blob = per layer frags | bias
// (padded to 8 NT), then the out scales [3].  Offsets must match HeadLayout in heads_mma.cu.
std::vector<float> pack_head_mma(const std::vector<HeadLayerSpec>& layers) {
    size_t total = 0;
    for (const HeadLayerSpec& L : layers) total += (size_t)mma_frag_words((L.K + 15) / 16, (L.N + 7) / 8) + (size_t)(L.N + 7) / 8 * 8;
    const size_t os = total;
    total = (total + 3 + 3) / 4 * 4;
    std::vector<float> v(total, 0.f);
    size_t off = 0;
    for (size_t l = 0; l < layers.size(); ++l) {
        const HeadLayerSpec& L = layers[l];
        const int KS = (L.K + 15) / 16, NT = (L.N + 7) / 8;
        v[os + l] = pack_mma_weight(L.W, L.K, iota_map(L.K, KS * 16), iota_map(L.N, NT * 8), &v[off]);
        off += (size_t)mma_frag_words(KS, NT);
        for (int o = 0; o < L.N; ++o) v[off + o] = L.b[o];
        off += (size_t)NT * 8;
    }
    return v;
}

std::vector<float> pack_data_bn(const StateDict& sd, const std::string& gp) {
    BnAffine a = bn_affine(sd, gp + "data_bn", 45);
    std::vector<float> v(90);
    for (int i = 0; i < 45; ++i) { v[i] = a.s[i]; v[45 + i] = a.o[i]; }
    return v;
}

// One st_gcn block (Net/GCN.py:67-147) as two GEMMs on channel-last rows (f*15 + joint):
//   gconv: U = relu( [Y Ahat_0 | Y Ahat_1] Wg^T + bias_w )                 K = 2*Cin, bias depends on the joint w
//          Wg[c'][k*Cin + c] = s_a[c'] conv.weight[k*C' + c'][c]
//          bias_w[w][c'] = s_a[c'] sum_k conv.bias[k*C'+c'] colsum_k[w] + o_a[c'],  colsum_k[w] = sum_v Ahat_k[v][w]
//   tconv: Y' = relu( sum_tau U(l+tau-4) Wt_tau^T + Y Wr^T + bias )         K = 9*C' + Cin
//          Wt_tau[c'][c] = s_b[c'] tcn.2.weight[c'][c][tau];  Wr = s_r residual.0.weight
//          bias = s_b tcn.2.bias + o_b + s_r residual.0.bias + o_r
HostGcnLayer pack_gcn_layer(const StateDict& sd, const std::string& gp, int layer, int Cin, int Cout) {
    HostGcnLayer L;
    L.cin = Cin;
    L.cout = Cout;
    const int V = kGcnV;
    const float* A = sd.get(gp + "A", 2 * V * V);
    const float* imp = sd.get(gp + "edge_importance." + std::to_string(layer), 2 * V * V);
    L.ahat.resize(2 * V * V);
    for (int i = 0; i < 2 * V * V; ++i) L.ahat[i] = A[i] * imp[i];
    const std::string g = gp + "gcn_networks." + std::to_string(layer) + ".";
    const float* cw = sd.get(g + "gcn.conv.weight", (long long)2 * Cout * Cin);
    const float* cb = sd.get(g + "gcn.conv.bias", 2 * Cout);
    BnAffine a = bn_affine(sd, g + "tcn.0", Cout);
    // gconv
    HostPackedGemm& gc = L.gconv;
    gc.N = Cout;
    gc.nseg = 1;
    gc.k[0] = 2 * Cin; gc.kpad[0] = pad16(2 * Cin);
    gc.ldw = gc.kpad[0];
    gc.w.assign((size_t)Cout * gc.ldw, 0.f);
    gc.bias.assign((size_t)V * Cout, 0.f);
    for (int o = 0; o < Cout; ++o)
        for (int k = 0; k < 2; ++k)
            for (int c = 0; c < Cin; ++c)
                gc.w[(size_t)o * gc.ldw + k * Cin + c] = a.s[o] * cw[(size_t)(k * Cout + o) * Cin + c];
    for (int w = 0; w < V; ++w) {
        float cs[2] = {0.f, 0.f};
        for (int k = 0; k < 2; ++k)
            for (int v = 0; v < V; ++v) cs[k] += L.ahat[k * V * V + v * V + w];
        for (int o = 0; o < Cout; ++o)
            gc.bias[(size_t)w * Cout + o] = a.s[o] * (cb[o] * cs[0] + cb[Cout + o] * cs[1]) + a.o[o];
    }
    // tconv
    const float* tw = sd.get(g + "tcn.2.weight", (long long)Cout * Cout * 9);
    const float* tb = sd.get(g + "tcn.2.bias", Cout);
    BnAffine b = bn_affine(sd, g + "tcn.3", Cout);
    const float* rw = sd.get(g + "residual.0.weight", (long long)Cout * Cin);
    const float* rb = sd.get(g + "residual.0.bias", Cout);
    BnAffine r = bn_affine(sd, g + "residual.1", Cout);
    HostPackedGemm& tc = L.tconv;
    tc.N = Cout;
    tc.nseg = 10;
    int kp = 0;
    for (int s = 0; s < 9; ++s) { tc.k[s] = Cout; tc.kpad[s] = pad16(Cout); kp += tc.kpad[s]; }
    tc.k[9] = Cin; tc.kpad[9] = pad16(Cin); kp += tc.kpad[9];
    tc.ldw = kp;
    tc.w.assign((size_t)Cout * kp, 0.f);
    tc.bias.assign(Cout, 0.f);
    for (int o = 0; o < Cout; ++o) {
        for (int tau = 0; tau < 9; ++tau)
            for (int c = 0; c < Cout; ++c)
                tc.w[(size_t)o * kp + tau * tc.kpad[0] + c] = b.s[o] * tw[((size_t)o * Cout + c) * 9 + tau];
        for (int c = 0; c < Cin; ++c) tc.w[(size_t)o * kp + 9 * tc.kpad[0] + c] = r.s[o] * rw[(size_t)o * Cin + c];
        tc.bias[o] = b.s[o] * tb[o] + b.o[o] + r.s[o] * rb[o] + r.o[o];
    }
    return L;
}

}  // namespace mmego
