// api.cu -- the C ABI of libmmego_b200 (include/mmego_b200.h): handle, weight upload, workspace planning and the
// per-stage kernel schedules.  Everything here is host code; the kernels live in the other .cu files.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/mmego_b200.h"
#include "internal.h"
#include "pack.h"

namespace mmego {
thread_local long long t_launches = 0;

// frees a device allocation owned by the handle (re-packing weights replaces buffers; nothing may pile up until destroy)
void handle_free(mmego_handle* h, void* p) {
    if (!p) return;
    for (size_t i = 0; i < h->owned.size(); ++i)
        if (h->owned[i] == p) {
            h->owned[i] = h->owned.back();
            h->owned.pop_back();
            break;
        }
    cudaFree(p);
}
}

using namespace mmego;

namespace {

std::string g_create_err;

int fail(mmego_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_err = buf;
    return code;
}

#define CUDA_TRY(h, expr)                                                                                \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) return fail((h), MMEGO_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// Per-call bookkeeping of an API entry point: selects the handle's device for the duration of the call (a process that
// drives several GPUs may have another one current) and credits the kernels launched by THIS call to THIS handle.
// Entry points nest (pipeline_forward -> imu_forward ...): only the outermost one counts.
struct Entry {
    mmego_handle* h;
    long long t0;
    int prev_dev = -1;
    bool outer = false;
    explicit Entry(mmego_handle* h_) : h(h_), t0(t_launches) {
        if (!h) return;
        outer = h->entry_depth++ == 0;
        int cur = -1;
        if (outer && cudaGetDevice(&cur) == cudaSuccess && cur != h->device) {
            prev_dev = cur;
            cudaSetDevice(h->device);
        }
    }
    ~Entry() {
        if (!h) return;
        --h->entry_depth;
        if (outer) h->launches += t_launches - t0;
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
    }
};

// ---------------------------------------------------------------------------------------------- device buffers
bool upload(mmego_handle* h, const std::vector<float>& v, DevBuf& out) {
    if (out.p && out.n != v.size()) {   // re-pack with a different size: release the old buffer
        handle_free(h, out.p);
        out.p = nullptr;
    }
    if (!out.p) {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(v.size(), 4) * sizeof(float)) != cudaSuccess) return false;
        h->owned.push_back(p);
        out.p = static_cast<float*>(p);
        out.n = v.size();
    }
    return cudaMemcpy(out.p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
}
bool upload(mmego_handle* h, const HostPackedGemm& g, PackedGemm& out) {
    out.N = g.N; out.ldw = g.ldw; out.nseg = g.nseg;
    for (int i = 0; i < kMaxSeg; ++i) { out.k[i] = g.k[i]; out.kpad[i] = g.kpad[i]; }
    return upload(h, g.w, out.w) && upload(h, g.bias, out.bias);
}

// ---------------------------------------------------------------------------------------------- workspace carving
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<char*>(b)) {}
    float* f(size_t n) {
        off = (off + 255) & ~size_t(255);
        float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
        off += n * sizeof(float);
        return p;
    }
};

#ifdef MMEGO_FFMA_GEN
// ---------------------------------------------------------------------------------------------- GEMM helpers (fp32 FFMA generation)
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

GemmArgs gemm_begin(const PackedGemm& w, float* c, long long ldc, long long M, int relu) {
    GemmArgs g{};
    g.nseg = 0;
    g.w = w.w.p;
    g.ldw = w.ldw;
    g.ktot = 0;
    g.bias = w.bias.p;
    g.c = c;
    g.ldc = ldc;
    g.M = (int)M;
    g.N = w.N;
    g.relu = relu;
    g.rowmod = 0;
    g.cstate = nullptr;
    g.ldcs = 0;
    g.has_state = 0;
    g.f6_period = 0;
    return g;
}
// appends the next packed K segment of `w` (segments must be added in packing order)
void gemm_seg(GemmArgs& g, const PackedGemm& w, const float* a, long long lda, int shift = 0, int period = 0) {
    const int s = g.nseg++;
    g.seg[s].a = a;
    g.seg[s].lda = lda;
    g.seg[s].k = w.k[s];
    g.seg[s].kpad = w.kpad[s];
    g.seg[s].shift = shift;
    g.seg[s].period = period;
    g.seg[s].vec = (aligned16(a) && lda % 4 == 0 && w.k[s] % 8 == 0) ? 1 : 0;
    g.ktot += w.kpad[s];
}
int pick_bn(int N) { return N >= 96 ? 128 : (N >= 48 ? 64 : 32); }

void run_gemm(mmego_handle* h, const GemmArgs& g, int epi, cudaStream_t st, int bn = 0) {
    GemmBatch b{};
    b.g[0] = g;
    launch_gemm(b, 1, bn ? bn : pick_bn(g.N), epi, st);
    (void)h;
}
// simple single-segment linear layer: c [M,N] = act(a [M,K] W^T + b)
void linear(mmego_handle* h, const PackedGemm& w, const float* a, long long lda, float* c, long long ldc, long long M,
            int relu, cudaStream_t st) {
    GemmArgs g = gemm_begin(w, c, ldc, M, relu);
    gemm_seg(g, w, a, lda);
    run_gemm(h, g, EPI_STORE, st);
}

#endif  // MMEGO_FFMA_GEN

void tap(mmego_handle* h, const char* name, const void* src, size_t bytes, cudaStream_t st) {
    auto it = h->taps.find(name);
    if (it == h->taps.end()) return;
    const size_t n = bytes < it->second.second ? bytes : it->second.second;
    cudaMemcpyAsync(it->second.first, src, n, cudaMemcpyDeviceToDevice, st);
    h->taps.erase(it);
}

// CUDA-event span around a group of launches on the launching stream (enabled by mmego_profile_begin)
struct Prof {
    mmego_handle* h;
    cudaStream_t st;
    size_t idx = (size_t)-1;
    long long l0 = 0;
    Prof(mmego_handle* h_, const char* name, cudaStream_t st_) : h(h_), st(st_) {
        if (!h->prof_on) return;
        ProfSpan sp;
        sp.name = name;
        if (cudaEventCreate(&sp.e0) != cudaSuccess || cudaEventCreate(&sp.e1) != cudaSuccess) return;
        cudaEventRecord(sp.e0, st);
        l0 = t_launches;
        h->prof.push_back(sp);
        idx = h->prof.size() - 1;
    }
    ~Prof() {
        if (idx == (size_t)-1) return;
        cudaEventRecord(h->prof[idx].e1, st);
        h->prof[idx].launches = t_launches - l0;
    }
};

// ---------------------------------------------------------------------------------------------- H=64 bi-LSTM stack
struct SmallLstmWs {
    float* gx;       // [S*T, 512]
    float* y[2];     // ping-pong [S*T, 128]
};
void plan_small_lstm(Carver& c, long long S, int T, SmallLstmWs& w) {
    w.gx = c.f((size_t)S * T * 512);
    w.y[0] = c.f((size_t)S * T * 128);
    w.y[1] = c.f((size_t)S * T * 128);
}
// in [S*T, In] -> returns pointer to the last layer's output [S*T, 128]
const float* run_small_lstm(mmego_handle* h, const PackedSmallLstmLayer* layers, const float* in, int in_ld,
                            const float* h0, const float* c0, float* hn, float* cn, long long S, int T,
                            const SmallLstmWs& w, cudaStream_t st) {
    const float* cur = in;
    long long ld = in_ld;
    Prof prof(h, "small_lstm", st);
    for (int l = 0; l < 3; ++l) {
#ifdef MMEGO_FFMA_GEN
        if (h->small_lstm_gemm)
#endif
        {
            const size_t so = (size_t)l * 2 * S * kSmallH;
            launch_lstm_small_mma(cur, ld, layers[l].in, layers[l].mma.p, w.gx, h0 ? h0 + so : nullptr,
                                  c0 ? c0 + so : nullptr, w.y[l & 1], hn ? hn + so : nullptr, cn ? cn + so : nullptr,
                                  (int)S, T, h->sm_count, st);
            cur = w.y[l & 1];
            ld = 128;
            const char* gxn[3] = {"small_lstm.gx0", "small_lstm.gx1", "small_lstm.gx2"};
            const char* yn[3] = {"small_lstm.y0", "small_lstm.y1", "small_lstm.y2"};
            tap(h, gxn[l], w.gx, (size_t)S * T * 512 * 4, st);
            tap(h, yn[l], cur, (size_t)S * T * 128 * 4, st);
            continue;
        }
#ifdef MMEGO_FFMA_GEN
        linear(h, layers[l].ih, cur, ld, w.gx, 512, S * T, 0, st);
        const size_t so = (size_t)l * 2 * S * kSmallH;
        launch_lstm_small(w.gx, layers[l].whh.p, h0 ? h0 + so : nullptr, c0 ? c0 + so : nullptr, w.y[l & 1],
                          hn ? hn + so : nullptr, cn ? cn + so : nullptr, (int)S, T, st);
        cur = w.y[l & 1];
        ld = 128;
#endif
    }
    return cur;
}

#ifdef MMEGO_FFMA_GEN
// ---------------------------------------------------------------------------------------------- IMU_Net schedule (fp32 FFMA generation)
struct ImuWs {
    float *u, *y0, *y1, *cst, *s, *z0, *z1;
};
void plan_imu(Carver& c, long long Bc, int L, int n, ImuWs& w) {
    const size_t S = (size_t)Bc * L;
    w.u = c.f(S * n * kImuH);
    w.y0 = c.f(S * n * 2 * kImuH);
    w.y1 = c.f(S * n * 2 * kImuH);
    w.cst = c.f(2 * S * kImuH);
    w.s = c.f(S * 2 * kImuH);
    w.z0 = c.f(S * 2 * kImuH);
    w.z1 = c.f(S * 2 * kImuH);
}

// one bidirectional H=512 layer over `T` steps for `S` sequences; x [S, T, In] -> y [S, T, 1024]
// Each timestep is ONE launch covering both directions (blockIdx.z) of an fp32 GEMM over K = [x_t | h_{t-1}] with the
// LSTM cell fused into the epilogue (gemm_ffma.cu).
void run_big_lstm_layer(mmego_handle* h, const PackedBigLstmLayer& lw, const float* x, int In, float* y, float* cst,
                        long long S, int T, cudaStream_t st, const char* span) {
    const int H = kImuH;
    Prof prof(h, span, st);
    for (int step = 0; step < T; ++step) {
        GemmBatch b{};
        for (int d = 0; d < 2; ++d) {
            const int tt = d ? (T - 1 - step) : step;
            const int tp = d ? (tt + 1) : (tt - 1);
            const PackedGemm& w = lw.dir[d];
            GemmArgs g = gemm_begin(w, y + (size_t)tt * 2 * H + d * H, (long long)T * 2 * H, S, 0);
            gemm_seg(g, w, x + (size_t)tt * In, (long long)T * In);
            if (step > 0) gemm_seg(g, w, y + (size_t)tp * 2 * H + d * H, (long long)T * 2 * H);
            g.cstate = cst + (size_t)d * S * H;
            g.ldcs = H;
            g.has_state = step > 0;
            b.g[d] = g;
        }
        launch_gemm(b, 2, 128, EPI_LSTM, st);
    }
    (void)h;
}

int imu_chunk_forward(mmego_handle* h, const float* imu, float* R, float* t, long long Bc, int L, int n,
                      const ImuWs& w, cudaStream_t st) {
    const long long S = Bc * L;
    const ImuWeights& W = h->imu;
    {
        Prof p(h, "imu.fc1", st);
        linear(h, W.fc1, imu, kImuFeat, w.u, kImuH, S * n, 1, st);                      // Net/IMU_Net.py:79
    }
    tap(h, "imu.u", w.u, (size_t)S * n * kImuH * 4, st);
    run_big_lstm_layer(h, W.fast[0], w.u, kImuH, w.y0, w.cst, S, n, st, "imu.lstm_fast");                 // :80
    run_big_lstm_layer(h, W.fast[1], w.y0, 2 * kImuH, w.y1, w.cst, S, n, st, "imu.lstm_fast");
    tap(h, "imu.f", w.y1, (size_t)S * n * 2 * kImuH * 4, st);
    {
        Prof p(h, "imu.pool", st);
        launch_imu_pool(w.y1, W.attn.p, w.s, S, n, st);                                  // :82-83
    }
    tap(h, "imu.s", w.s, (size_t)S * 2 * kImuH * 4, st);
    run_big_lstm_layer(h, W.slow[0], w.s, 2 * kImuH, w.z0, w.cst, Bc, L, st, "imu.lstm_slow");            // :85
    run_big_lstm_layer(h, W.slow[1], w.z0, 2 * kImuH, w.z1, w.cst, Bc, L, st, "imu.lstm_slow");
    tap(h, "imu.g", w.z1, (size_t)S * 2 * kImuH * 4, st);
    {
        Prof p(h, "imu.decode", st);
        launch_imu_decode(w.z1, W.fc2.p, R, t, S, st);                                   // :87-93
    }
    return MMEGO_OK;
}

#endif  // MMEGO_FFMA_GEN

// ---------------------------------------------------------------------------------------------- IMU_Net latency path
// Small batches (B*L <= kResMaxSeq): fp32 throughout, persistent LSTM kernels with the gate weights resident in shared
// memory (lstm_resident.cu).  7 launches per call instead of 83.
struct ImuResWs {
    float *u, *y0, *y1, *s, *z0, *z1, *cst, *gxs;
    unsigned* flags;
    unsigned long long* xchg;
};
void plan_imu_res(Carver& c, long long B, int L, int n, ImuResWs& w) {
    const size_t S = (size_t)B * L;
    w.u = c.f(S * n * kImuH);
    w.y0 = c.f(S * n * 2 * kImuH);
    w.y1 = c.f(S * n * 2 * kImuH);
    w.s = c.f(S * 2 * kImuH);
    w.z0 = c.f(S * 2 * kImuH);
    w.z1 = c.f(S * 2 * kImuH);
    w.cst = c.f(2 * S * kImuH);
    w.flags = reinterpret_cast<unsigned*>(c.f(2 * (size_t)(n > L ? n : L) + 8));
    w.gxs = c.f(resident_gx_floats((int)B, L));      // rnn_slow: B sequences x L timesteps
    w.xchg = reinterpret_cast<unsigned long long*>(c.f(2 * resident_xchg_words((int)S)));       // S >= B
}
bool use_resident(const mmego_handle* h, long long B, int L) {
    return h->imu_resident && h->imu.res_ready && B * L <= h->imu_res_max_seq && resident_supported(h->sm_count);
}
int imu_res_forward(mmego_handle* h, const float* imu, float* R, float* t, long long B, int L, int n, const ImuResWs& w,
                    cudaStream_t st) {
    const long long S = B * L;
    const ImuWeights& W = h->imu;
    Prof p(h, "imu.resident", st);
    launch_res_fc1(imu, W.res_fc1.p, w.u, S * n, st);                                                            // Net/IMU_Net.py:79
    int rc = 0;
    float* const gxs = h->imu_res_pre ? w.gxs : nullptr;
    unsigned long long* const xc = h->imu_res_xchg ? w.xchg : nullptr;
    // rnn_fast (S = B*L sequences of n samples): tensor-core form; rnn_slow (B sequences of L frames): fp32 with its input
    // projections up front.  The choice never depends on the batch size, so results do not depend on batch-mates / chunking.
    const bool tc = h->imu_res_tc && W.res_wtc[0].p && W.res_wtc[1].p;
    rc |= launch_lstm_resident(w.u, kImuH, w.y0, tc ? W.res_wtc[0].p : W.res_w[0].p, tc ? W.res_stc[0].p : nullptr,
                               W.res_b[0].p, w.cst, w.flags, nullptr, xc, h->imu_res_direct, (int)S, n, st);                                    // :80
    rc |= launch_lstm_resident(w.y0, 2 * kImuH, w.y1, tc ? W.res_wtc[1].p : W.res_w[1].p, tc ? W.res_stc[1].p : nullptr,
                               W.res_b[1].p, w.cst, w.flags, nullptr, xc, h->imu_res_direct, (int)S, n, st);
    tap(h, "imu.f", w.y1, (size_t)S * n * 2 * kImuH * 4, st);
    launch_imu_pool(w.y1, W.attn.p, w.s, S, n, st);                                                              // :82-83
    rc |= launch_lstm_resident(w.s, 2 * kImuH, w.z0, W.res_w[2].p, nullptr, W.res_b[2].p, w.cst, w.flags, gxs, xc, h->imu_res_direct, (int)B, L, st);  // :85
    rc |= launch_lstm_resident(w.z0, 2 * kImuH, w.z1, W.res_w[3].p, nullptr, W.res_b[3].p, w.cst, w.flags, gxs, xc, h->imu_res_direct, (int)B, L, st);
    launch_imu_decode(w.z1, W.fc2.p, R, t, S, st);                                                               // :87-93
    if (rc) return fail(h, MMEGO_ECUDA, "imu_forward: launching the resident-weights LSTM kernel failed (%s)",
                        cudaGetErrorString(cudaGetLastError()));
    return MMEGO_OK;
}

#ifndef MMEGO_EMUL
// ---------------------------------------------------------------------------------------------- IMU_Net on tensor cores
// Activations are fp16 hi/lo planes (see lstm_tc.cu); one plane of n elements takes n/2 floats of workspace.
struct ImuTcWs {
    void *u[2], *y0[2], *y1[2], *s[2], *z0[2], *z1[2];
    float* cst;
    long long Spad;
};
// rnn_fast works on chunks of Bc snippets (u, y0, y1: 4.3 GB per 2048 snippets); the pooled features s of ALL B snippets
// are kept (80 KB per snippet) so that rnn_slow and the decode run ONCE over the whole batch after the chunk loop: half
// the rnn_slow launches at B = 4096, each with twice the work items for the 74 CTA pairs.
void plan_imu_tc(Carver& c, long long B, long long Bc, int L, int n, ImuTcWs& w) {
    const size_t S = (size_t)Bc * L, Sall = (size_t)B * L;
    auto plane = [&](size_t elems) { return static_cast<void*>(c.f((elems + 1) / 2)); };
    for (int k = 0; k < 2; ++k) {
        w.u[k] = plane(S * n * kImuH);
        w.y0[k] = plane(S * n * 2 * kImuH);
        w.y1[k] = plane(S * n * 2 * kImuH);
        w.s[k] = plane(Sall * 2 * kImuH);
        w.z0[k] = plane(Sall * 2 * kImuH);
        w.z1[k] = plane(Sall * 2 * kImuH);
    }
    const size_t seqs = S > (size_t)B ? S : (size_t)B;      // rnn_fast: Bc*L sequences per chunk; rnn_slow: B sequences
    w.Spad = (long long)((seqs + 127) / 128 * 128);
    w.cst = c.f((size_t)2 * kImuH * w.Spad);
}

// rnn_fast part of one chunk: imu [Bc,L,n,15] -> pooled features s (planes of the WHOLE batch, this chunk's rows at f0)
int imu_chunk_fast_tc(mmego_handle* h, const float* imu, long long f0, long long Bc, int L, int n, const ImuTcWs& w,
                      cudaStream_t st) {
    const long long S = Bc * L;
    const ImuWeights& W = h->imu;
    const int npass = h->imu_gemm == 1 ? 3 : 1;
    void* const nolo = nullptr;
    auto lo = [&](void* const* planes) { return npass == 3 ? planes[1] : nolo; };
    auto at = [&](void* plane, long long elems) { return plane ? static_cast<void*>(static_cast<char*>(plane) + elems * 2) : nolo; };
    {
        Prof p(h, "imu.fc1", st);
        tc_imu_fc1(imu, W.fc1_mma.p, w.u[0], lo(w.u), S * n, h->sm_count, h->tc_lo_drop, st);                               // Net/IMU_Net.py:79
    }
    auto tap_split = [&](const char* name, void* const* planes, long long elems) {
        auto it = h->taps.find(name);
        if (it == h->taps.end()) return;
        tc_unsplit(planes[0], lo(planes), static_cast<float*>(it->second.first),
                   std::min<long long>(elems, (long long)(it->second.second / 4)), st);
        h->taps.erase(it);
    };
    tap_split("imu.u", w.u, S * n * kImuH);
    int rc = 0;
    {
        Prof p(h, "imu.lstm_fast", st);
        rc |= tc_lstm_layer(h, W.tc_fast[0], w.u[0], lo(w.u), w.y0[0], lo(w.y0), w.cst, S, w.Spad, n, npass, (h->tc_persist & 2) != 0, st);   // :80
        tap_split("imu.y0", w.y0, S * n * 2 * kImuH);
        rc |= tc_lstm_layer(h, W.tc_fast[1], w.y0[0], lo(w.y0), w.y1[0], lo(w.y1), w.cst, S, w.Spad, n, npass, (h->tc_persist & 2) != 0, st);
    }
    tap_split("imu.f", w.y1, S * n * 2 * kImuH);
    {
        Prof p(h, "imu.pool", st);
        tc_imu_pool(w.y1[0], lo(w.y1), W.attn.p, at(w.s[0], f0 * 2 * kImuH), at(lo(w.s), f0 * 2 * kImuH), S, n, h->tc_lo_drop, st);   // :82-83
    }
    if (rc) return fail(h, MMEGO_ECUDA, "imu_forward: cuTensorMapEncodeTiled failed");
    return MMEGO_OK;
}

// rnn_slow + fc2 + 6D decode over the whole batch: s planes [B*L,1024] -> R [B,L,3,3], t [B,L,3]
int imu_slow_decode_tc(mmego_handle* h, float* R, float* t, long long B, int L, const ImuTcWs& w, cudaStream_t st) {
    const long long S = B * L;
    const ImuWeights& W = h->imu;
    const int npass = h->imu_gemm == 1 ? 3 : 1;
    void* const nolo = nullptr;
    auto lo = [&](void* const* planes) { return npass == 3 ? planes[1] : nolo; };
    auto tap_split = [&](const char* name, void* const* planes, long long elems) {
        auto it = h->taps.find(name);
        if (it == h->taps.end()) return;
        tc_unsplit(planes[0], lo(planes), static_cast<float*>(it->second.first),
                   std::min<long long>(elems, (long long)(it->second.second / 4)), st);
        h->taps.erase(it);
    };
    tap_split("imu.s", w.s, S * 2 * kImuH);
    int rc = 0;
    {
        Prof p(h, "imu.lstm_slow", st);
        rc |= tc_lstm_layer(h, W.tc_slow[0], w.s[0], lo(w.s), w.z0[0], lo(w.z0), w.cst, B, w.Spad, L, npass, (h->tc_persist & 1) != 0, st);  // :85
        rc |= tc_lstm_layer(h, W.tc_slow[1], w.z0[0], lo(w.z0), w.z1[0], lo(w.z1), w.cst, B, w.Spad, L, npass, (h->tc_persist & 1) != 0, st);
    }
    tap_split("imu.g", w.z1, S * 2 * kImuH);
    {
        Prof p(h, "imu.decode", st);
        tc_imu_decode(w.z1[0], lo(w.z1), W.fc2.p, R, t, S, st);                           // :87-93
    }
    if (rc) return fail(h, MMEGO_ECUDA, "imu_forward: cuTensorMapEncodeTiled failed");
    return MMEGO_OK;
}
#endif

// ---------------------------------------------------------------------------------------------- Upper_Net schedule
struct UpperWs {
    float* g;        // [F,64]
    SmallLstmWs lstm;
    float* h1;       // [F,128]
    float* o;        // [F,87]
};
void plan_upper(Carver& c, long long B, int L, UpperWs& w) {
    const size_t F = (size_t)B * L;
    w.g = c.f(F * 64);
    plan_small_lstm(c, B, L, w.lstm);
    w.h1 = c.f(F * 128);
    w.o = c.f(F * 87);
}

// ---------------------------------------------------------------------------------------------- Lower_Net schedule
struct LowerWs {
    float *uh, *ya, *u, *y[2], *kf, *ak, *f0, *f1, *o;
    void *p_y0[2], *p_ya[2], *p_u[2], *p_y[2][2];     // tensor-core path: fp16 hi/lo planes
    SmallLstmWs lstm;
};
void plan_lower(Carver& c, long long B, int L, LowerWs& w, bool tc) {
    const size_t F = (size_t)B * L, FV = F * kGcnV;
    w.uh = c.f(F * 45);
    if (tc) {
        auto plane = [&](size_t elems) { return static_cast<void*>(c.f((elems + 1) / 2)); };
        for (int k = 0; k < 2; ++k) {
            w.p_y0[k] = plane(FV * 8);
            w.p_ya[k] = plane(FV * 128);
            w.p_u[k] = plane(FV * 128);
            w.p_y[0][k] = plane(FV * 128);
            w.p_y[1][k] = plane(FV * 128);
        }
        w.ya = w.u = w.y[0] = w.y[1] = nullptr;
    } else {
        w.ya = c.f(FV * 128);       // aggregated input of the widest layer: 2 * 64
        w.u = c.f(FV * 128);        // graph-conv output of the widest layer
        w.y[0] = c.f(FV * 128);
        w.y[1] = c.f(FV * 128);
    }
    w.kf = c.f(FV * 64);
    w.ak = c.f(F * 192);
    plan_small_lstm(c, B, L, w.lstm);
    w.f0 = c.f(F * 128);
    w.f1 = c.f(F * 64);
    w.o = c.f(F * 42);
}

#ifdef MMEGO_FFMA_GEN
// GCN.Model.extract_feature body on channel-last rows; y0 [F*15, 3] (data_bn already applied) -> kf [B][64][L*15]
void run_gcn(mmego_handle* h, const float* y0, int B, int L, const LowerWs& w, cudaStream_t st) {
    const LowerWeights& W = h->lower;
    const long long F = (long long)B * L, FV = F * kGcnV;
    const float* y = y0;
    for (int i = 0; i < 3; ++i) {
        const GcnLayerWeights& g = W.gcn[i];
        launch_gcn_agg(y, g.ahat.p, w.ya, F, g.cin, h->sm_count, st);                    // Net/GCN.py:62 (commuted)
        {
            GemmArgs a = gemm_begin(g.gconv, w.u, g.cout, FV, 1);                        // :58 + tcn.0/1 (BN, ReLU)
            gemm_seg(a, g.gconv, w.ya, 2 * g.cin);
            a.rowmod = kGcnV;
            run_gemm(h, a, EPI_STORE, st);
        }
        {
            float* out = w.y[i & 1];
            GemmArgs a = gemm_begin(g.tconv, out, g.cout, FV, 1);                        // :109-116 + residual :128-147
            for (int tau = 0; tau < 9; ++tau) gemm_seg(a, g.tconv, w.u, g.cout, (tau - 4) * kGcnV, L * kGcnV);
            gemm_seg(a, g.tconv, y, g.cin);
            run_gemm(h, a, EPI_STORE, st);
            y = out;
        }
    }
    GemmArgs a = gemm_begin(W.fcn, w.kf, 0, FV, 0);                                      // :352-353 (F6 layout)
    gemm_seg(a, W.fcn, y, 128);
    a.f6_period = L * kGcnV;
    run_gemm(h, a, EPI_F6, st, 64);
}
#endif  // MMEGO_FFMA_GEN

#ifndef MMEGO_EMUL
// ST-GCN on tensor cores: y0 planes [F*15][8] (data_bn applied) -> kf fp32 [B][64][L*15]
int run_gcn_tc(mmego_handle* h, int B, int L, const LowerWs& w, cudaStream_t st) {
    const LowerWeights& W = h->lower;
    const long long F = (long long)B * L, FV = F * kGcnV;
    const int RP = L * kGcnV;
    void* const* y = w.p_y0;
    int ystride = 8, creal = 3;
    int rc = 0;
    for (int i = 0; i < 3; ++i) {
        const int cout = W.gcn[i].cout;
        const int os = i == 0 ? 8 : 2 * creal;
        if (i == 0) {   // channels 6, 7 of the 8-wide aggregated rows are padding
            cudaMemsetAsync(w.p_ya[0], 0, (size_t)FV * 8 * 2, st);
            cudaMemsetAsync(w.p_ya[1], 0, (size_t)FV * 8 * 2, st);
        }
        static const char* kAgg[3] = {"gcn.agg0", "gcn.agg1", "gcn.agg2"};
        static const char* kGc[3] = {"gcn.gconv0", "gcn.gconv1", "gcn.gconv2"};
        static const char* kTc[3] = {"gcn.tconv0", "gcn.tconv1", "gcn.tconv2"};
        {
            Prof p(h, kAgg[i], st);
            tc_gcn_agg(y[0], y[1], W.gcn[i].ahat.p, w.p_ya[0], w.p_ya[1], F, creal, ystride, os, h->sm_count, st);
        }
        {
            Prof p(h, kGc[i], st);
            rc |= tc_gcn_gemm(h, W.tc_gconv[i], w.p_ya[0], w.p_ya[1], os, 1, nullptr, nullptr, 0, kGcnV, 1, w.p_u[0], w.p_u[1],
                              nullptr, B, RP, st);
        }
        void* const* out = w.p_y[i & 1];
        {
            Prof p(h, kTc[i], st);
            // gcn_snip: bit 0 on; bit 4 / bit 12+i a second drain group per block (all layers / layer i); bit 8+i: layer i
            // on the row-tiled kernel
            if ((h->gcn_snip & 1) && !(h->gcn_snip & (256 << i)) && tc_gcn_tconv_snip_supported(L))
                rc |= tc_gcn_tconv_snip(h, W.tc_tconv[i], w.p_u[0], w.p_u[1], cout, y[0], y[1], ystride, creal, out[0], out[1], B, L,
                                        (h->gcn_snip & (16 | (4096 << i))) != 0, st);
            else
                rc |= tc_gcn_gemm(h, W.tc_tconv[i], w.p_u[0], w.p_u[1], cout, 9, y[0], y[1], ystride, 0, 1, out[0], out[1],
                                  nullptr, B, RP, st);
        }
        y = out;
        ystride = creal = cout;
    }
    {
        Prof p(h, "gcn.fcn", st);
        rc |= tc_gcn_gemm(h, W.tc_fcn, y[0], y[1], 128, 1, nullptr, nullptr, 0, 0, 0, nullptr, nullptr, w.kf, B, RP, st);
    }
    return rc;
}
#endif

bool use_gcn_tc(const mmego_handle* h) {
#ifdef MMEGO_EMUL
    (void)h;
    return false;
#else
    return h->gcn_gemm != 0;
#endif
}

size_t imu_ws_bytes(const mmego_handle* h, long long B, long long Bc, int L, int n) {
    Carver s(nullptr);
    if (use_resident(h, B, L)) {
        ImuResWs w;
        plan_imu_res(s, B, L, n, w);
        return s.off;
    }
#ifndef MMEGO_EMUL
    if (h->imu_gemm != 0) {
        ImuTcWs w;
        plan_imu_tc(s, B, Bc, L, n, w);
        return s.off;
    }
#endif
#ifdef MMEGO_FFMA_GEN
    ImuWs w;
    plan_imu(s, Bc, L, n, w);
#endif
    (void)h;
    return s.off;
}

int check_dims(mmego_handle* h, int B, int L, int N) {
    if (!h) return MMEGO_EINVAL;
    if (B <= 0 || L <= 0 || N <= 0) return fail(h, MMEGO_ESHAPE, "B, L, N must be positive (got %d, %d, %d)", B, L, N);
    if ((long long)B * L * kGcnV * 128 > 2000000000LL) return fail(h, MMEGO_ESHAPE, "batch too large for one call: B*L = %lld", (long long)B * L);
    return MMEGO_OK;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int mmego_abi_version(void) { return MMEGO_ABI_VERSION; }

int mmego_create(mmego_handle** out, int device) {
    if (!out) return fail(nullptr, MMEGO_EINVAL, "mmego_create: out is NULL");
    *out = nullptr;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, MMEGO_ECUDA, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, MMEGO_EARCH, "device %d (%s) is sm_%d%d; libmmego_b200 is built for sm_100a only", device,
                    prop.name, prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, MMEGO_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    mmego_handle* h = new mmego_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
#ifdef MMEGO_EMUL
    h->imu_gemm = 0;
#else
    h->imu_gemm = tc_supported() ? 1 : 0;      // default: tcgen05 fp16x3 (fp32-grade); 2 = tcgen05 fp16; 0 = fp32 FFMA (test builds)
    h->gcn_gemm = tc_supported() ? 1 : 0;
#ifndef MMEGO_FFMA_GEN
    if (!tc_supported()) {
        delete h;
        return fail(nullptr, MMEGO_EARCH, "cuTensorMapEncodeTiled is not available from this driver: the tensor-core kernels cannot run (there is no fallback path)");
    }
#endif
#endif
    *out = h;
    return MMEGO_OK;
}

int mmego_destroy(mmego_handle* h) {
    if (!h) return MMEGO_EINVAL;
    cudaSetDevice(h->device);
    for (void* p : h->owned) cudaFree(p);
    if (h->stage_dev) cudaFree(h->stage_dev);
    if (h->tc_stats) cudaFree(h->tc_stats);
    if (h->dev_error) cudaFree(h->dev_error);
    if (h->tc_sync) cudaFree(h->tc_sync);
    for (ProfSpan& sp : h->prof) {
        if (sp.e0) cudaEventDestroy(sp.e0);
        if (sp.e1) cudaEventDestroy(sp.e1);
    }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    for (cudaEvent_t e : h->host_events) cudaEventDestroy(e);
    delete h;
    return MMEGO_OK;
}

const char* mmego_last_error(const mmego_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int mmego_set_option(mmego_handle* h, const char* key, long long value) {
    Entry entry(h);
    if (!h || !key) return MMEGO_EINVAL;
    if (!strcmp(key, "imu_chunk")) {
        if (value <= 0) return fail(h, MMEGO_EINVAL, "imu_chunk must be positive");
        h->imu_chunk = value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "imu_gemm")) {
        if (value < 0 || value > 2) return fail(h, MMEGO_EINVAL, "imu_gemm must be 0 (fp32 FFMA), 1 (tcgen05 fp16x3) or 2 (tcgen05 fp16)");
#ifndef MMEGO_FFMA_GEN
        if (value == 0) return fail(h, MMEGO_EINVAL, "imu_gemm=0: the fp32 FFMA kernel generation is not part of the product library (test builds: -DMMEGO_WITH_FFMA)");
#endif
#ifdef MMEGO_EMUL
        if (value != 0) return fail(h, MMEGO_EINVAL, "imu_gemm=%lld needs the sm_100a build", value);
#else
        if (value != 0 && !tc_supported()) return fail(h, MMEGO_EARCH, "imu_gemm=%lld: cuTensorMapEncodeTiled is not available from this driver", value);
        if (value != 0 && h->imu.ready && !h->imu.tc_ready) return fail(h, MMEGO_ESTATE, "imu_gemm=%lld: tensor-core weight packing failed at set_weights", value);
#endif
        h->imu_gemm = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "tc_dbg")) {
#ifdef MMEGO_DEBUG_SWITCHES
        h->tc_dbg = (int)value;      // experiment switches of lstm_tc.cu (some make the kernel skip work): test builds only
        return MMEGO_OK;
#else
        return fail(h, MMEGO_EINVAL, "tc_dbg exists only in builds with -DMMEGO_DEBUG_SWITCHES (it is not part of the product library)");
#endif
    }
    if (!strcmp(key, "tc_cta_pair")) {
        h->tc_cta_pair = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "host_chunk")) {
        if (value <= 0) return fail(h, MMEGO_EINVAL, "host_chunk must be positive");
        h->host_chunk = value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "gcn_gemm")) {
        if (value < 0 || value > 1) return fail(h, MMEGO_EINVAL, "gcn_gemm must be 0 (fp32 FFMA) or 1 (tcgen05 fp16x3)");
#ifdef MMEGO_EMUL
        if (value != 0) return fail(h, MMEGO_EINVAL, "gcn_gemm=1 needs the sm_100a build");
#else
        if (value != 0 && !tc_supported()) return fail(h, MMEGO_EARCH, "gcn_gemm=1: cuTensorMapEncodeTiled is not available");
#endif
#ifndef MMEGO_FFMA_GEN
        if (value == 0) return fail(h, MMEGO_EINVAL, "gcn_gemm=0: the fp32 FFMA kernel generation is not part of the product library (test builds: -DMMEGO_WITH_FFMA)");
#endif
        h->gcn_gemm = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "point_gemm")) {
        if (value < 0 || value > 1) return fail(h, MMEGO_EINVAL, "point_gemm must be 0 (fp32 FFMA) or 1 (mma.sync fp16x3)");
#ifndef MMEGO_FFMA_GEN
        if (value == 0) return fail(h, MMEGO_EINVAL, "point_gemm=0: the fp32 FFMA kernel generation is not part of the product library (test builds: -DMMEGO_WITH_FFMA)");
#endif
        h->point_gemm = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "gcn_kb_chunk")) {
        if (value < 0 || value > 64) return fail(h, MMEGO_EINVAL, "gcn_kb_chunk must be in 0..64");
        h->gcn_kb_chunk = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "head_gemm")) {
        if (value < 0 || value > 1) return fail(h, MMEGO_EINVAL, "head_gemm must be 0 (fp32 FFMA) or 1 (mma.sync fp16x3)");
#ifndef MMEGO_FFMA_GEN
        if (value == 0) return fail(h, MMEGO_EINVAL, "head_gemm=0: the fp32 FFMA kernel generation is not part of the product library (test builds: -DMMEGO_WITH_FFMA)");
#endif
        h->head_gemm = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "small_lstm_gemm")) {
        if (value < 0 || value > 1) return fail(h, MMEGO_EINVAL, "small_lstm_gemm must be 0 (fp32 FFMA) or 1 (mma.sync fp16x3)");
#ifndef MMEGO_FFMA_GEN
        if (value == 0) return fail(h, MMEGO_EINVAL, "small_lstm_gemm=0: the fp32 FFMA kernel generation is not part of the product library (test builds: -DMMEGO_WITH_FFMA)");
#endif
        h->small_lstm_gemm = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "tc_lo_drop")) {
        if (value < 0 || value > 6) return fail(h, MMEGO_EINVAL, "tc_lo_drop must be in 0..6");
#ifndef MMEGO_EMUL
        // the packed weight planes are rounded in place: the option can only grow once weights are loaded
        if (h->imu.tc_ready && value < h->tc_lo_drop_w)
            return fail(h, MMEGO_ESTATE, "tc_lo_drop: the packed weights are already rounded to fewer bits; reload IMU_Net first");
        h->tc_lo_drop = (int)value;
        if (h->imu.tc_ready && value > h->tc_lo_drop_w) {
            cudaSetDevice(h->device);
            for (int l = 0; l < 2; ++l) {
                tc_round_lo_weights(h->imu.tc_fast[l], h->tc_lo_drop, nullptr);
                tc_round_lo_weights(h->imu.tc_slow[l], h->tc_lo_drop, nullptr);
            }
            if (cudaDeviceSynchronize() != cudaSuccess) return fail(h, MMEGO_ECUDA, "tc_lo_drop: rounding the weight planes failed");
            h->tc_lo_drop_w = h->tc_lo_drop;
        }
#else
        h->tc_lo_drop = (int)value;
#endif
        return MMEGO_OK;
    }
    if (!strcmp(key, "point_stage")) {
        h->point_stage = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "gcn_w_res")) {
        h->gcn_w_res = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "gcn_snip")) {
        h->gcn_snip = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "imu_res_max_seq")) {
        if (value < 0 || value > 4096) return fail(h, MMEGO_EINVAL, "imu_res_max_seq must be in 0..4096");
        h->imu_res_max_seq = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "imu_res_direct")) {
        h->imu_res_direct = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "imu_res_xchg")) {
        h->imu_res_xchg = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "imu_res_tc")) {
        h->imu_res_tc = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "imu_res_pre")) {
        h->imu_res_pre = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "imu_resident")) {
        h->imu_resident = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "tc_persist")) {
        if (value < 0 || value > 3) return fail(h, MMEGO_EINVAL, "tc_persist must be in 0..3");
        h->tc_persist = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "tc_pdl")) {
        h->tc_pdl = value != 0;
        return MMEGO_OK;
    }
    if (!strcmp(key, "tc_kb_chunk0")) {
        if (value < 0 || value > 64) return fail(h, MMEGO_EINVAL, "tc_kb_chunk0 must be in 0..64");
        h->tc_kb_chunk0 = (int)value;
        return MMEGO_OK;
    }
    if (!strcmp(key, "tc_kb_chunk")) {
        if (value < 0 || value > 64) return fail(h, MMEGO_EINVAL, "tc_kb_chunk must be in 0..64");
        h->tc_kb_chunk = (int)value;
        return MMEGO_OK;
    }
    return fail(h, MMEGO_EINVAL, "unknown option '%s'", key);
}

int mmego_set_weights(mmego_handle* h, int net, const char* const* names, const float* const* ptrs_host,
                      const long long* numels, int n) {
    Entry entry(h);
    if (!h || !names || !ptrs_host || !numels || n < 0) return MMEGO_EINVAL;
    cudaSetDevice(h->device);
    StateDict sd;
    for (int i = 0; i < n; ++i) {
        if (!names[i] || !ptrs_host[i]) return fail(h, MMEGO_EINVAL, "set_weights: entry %d is NULL", i);
        sd.m[names[i]] = {ptrs_host[i], numels[i]};
    }
    bool ok = true;
    try {
        if (net == MMEGO_NET_IMU) {
            ImuWeights& W = h->imu;
            W.ready = false;
            const int H = kImuH;
            {
                const HostPackedGemm fc1 = pack_linear(sd.get("fc1.weight", H * kImuFeat), sd.get("fc1.bias", H), H, {kImuFeat});
                ok &= upload(h, pack_imu_fc1_mma(fc1), W.fc1_mma);
#ifdef MMEGO_FFMA_GEN
                ok &= upload(h, fc1, W.fc1);
#endif
            }
#ifdef MMEGO_FFMA_GEN
            for (int l = 0; l < 2; ++l) {       // 92 MB of fp32 GEMM-format weights: only where the FFMA kernels exist
                HostBigLstm f = pack_big_lstm(sd, "rnn_fast.", l, l == 0 ? H : 2 * H, H);
                HostBigLstm s = pack_big_lstm(sd, "rnn_slow.", l, 2 * H, H);
                for (int d = 0; d < 2; ++d) {
                    ok &= upload(h, f.dir[d], W.fast[l].dir[d]);
                    ok &= upload(h, s.dir[d], W.slow[l].dir[d]);
                }
            }
#endif
            std::vector<float> attn(2 * H + 4, 0.f);
            memcpy(attn.data(), sd.get("attn.weight", 2 * H), sizeof(float) * 2 * H);
            attn[2 * H] = sd.get("attn.bias", 1)[0];
            ok &= upload(h, attn, W.attn);
            std::vector<float> fc2(9 * 2 * H + 12, 0.f);
            memcpy(fc2.data(), sd.get("fc2.weight", 9 * 2 * H), sizeof(float) * 9 * 2 * H);
            memcpy(fc2.data() + 9 * 2 * H, sd.get("fc2.bias", 9), sizeof(float) * 9);
            ok &= upload(h, fc2, W.fc2);
            {   // latency path: fp32 slices per (direction, 8-unit group), k-major (lstm_resident.cu)
                std::vector<float> rw, rb;
                const char* pre[4] = {"rnn_fast.", "rnn_fast.", "rnn_slow.", "rnn_slow."};
                for (int i = 0; i < 4; ++i) {
                    pack_resident_layer(sd, pre[i], i & 1, i == 0 ? H : 2 * H, rw, rb);
                    ok &= upload(h, rw, W.res_w[i]) && upload(h, rb, W.res_b[i]);
                    if (i < 2) {          // rnn_fast also as tensor-core fragments
                        pack_resident_layer_tc(sd, pre[i], i & 1, i == 0 ? H : 2 * H, rw, rb);
                        ok &= upload(h, rw, W.res_wtc[i]) && upload(h, rb, W.res_stc[i]);
                    }
                }
                std::vector<float> f1((size_t)H * kImuFeat + H);
                memcpy(f1.data(), sd.get("fc1.weight", H * kImuFeat), sizeof(float) * H * kImuFeat);
                memcpy(f1.data() + (size_t)H * kImuFeat, sd.get("fc1.bias", H), sizeof(float) * H);
                ok &= upload(h, f1, W.res_fc1);
                W.res_ready = ok;
            }
            W.ready = ok;
            W.tc_ready = false;
#ifndef MMEGO_EMUL
            if (ok && tc_supported()) {
                bool t = true;
                for (int l = 0; l < 2; ++l) {
                    t = t && tc_pack_layer(h, sd, "rnn_fast.", l, l == 0 ? H : 2 * H, W.tc_fast[l]);
                    t = t && tc_pack_layer(h, sd, "rnn_slow.", l, 2 * H, W.tc_slow[l]);
                }
                W.tc_ready = t;
                h->tc_lo_drop_w = 0;
                if (t && h->tc_lo_drop > 0) {
                    for (int l = 0; l < 2; ++l) {
                        tc_round_lo_weights(W.tc_fast[l], h->tc_lo_drop, nullptr);
                        tc_round_lo_weights(W.tc_slow[l], h->tc_lo_drop, nullptr);
                    }
                    if (cudaDeviceSynchronize() != cudaSuccess) return fail(h, MMEGO_ECUDA, "set_weights: rounding the weight planes failed");
                    h->tc_lo_drop_w = h->tc_lo_drop;
                }
            }
            if (h->imu_gemm != 0 && !W.tc_ready) return fail(h, MMEGO_ESTATE, "set_weights: tensor-core packing of IMU_Net failed");
#endif
        } else if (net == MMEGO_NET_UPPER) {
            UpperWeights& W = h->upper;
            W.ready = false;
            {
                const std::vector<float> folded = pack_upper_point(sd);
                ok &= upload(h, folded, W.point) && upload(h, pack_upper_point_mma(folded), W.point_mma);
            }
            for (int l = 0; l < 3; ++l) {
                HostSmallLstm s = pack_small_lstm(sd, "module1.grnn.", l, l == 0 ? 64 : 128);
                ok &= upload(h, s.ih, W.lstm[l].ih) && upload(h, s.whh, W.lstm[l].whh);
                ok &= upload(h, pack_small_lstm_mma(sd, "module1.grnn.", l, l == 0 ? 64 : 128), W.lstm[l].mma);
                W.lstm[l].in = s.in;
            }
            ok &= upload(h, pack_linear(sd.get("mlpHead.fc1.weight", 128 * 128), sd.get("mlpHead.fc1.bias", 128), 128, {128}), W.fc1);
            ok &= upload(h, pack_linear(sd.get("mlpHead.fc2.weight", 87 * 128), sd.get("mlpHead.fc2.bias", 87), 87, {128}), W.fc2);
            {
                const std::vector<float> hb = pack_head_mma({{sd.get("mlpHead.fc1.weight", 128 * 128), sd.get("mlpHead.fc1.bias", 128), 128, 128},
                                                             {sd.get("mlpHead.fc2.weight", 87 * 128), sd.get("mlpHead.fc2.bias", 87), 87, 128}});
                if (hb.size() != upper_head_mma_words()) return fail(h, MMEGO_ESHAPE, "set_weights: head layout mismatch");
                ok &= upload(h, hb, W.head_mma);
            }
            W.ready = ok;
        } else if (net == MMEGO_NET_LOWER) {
            LowerWeights& W = h->lower;
            W.ready = false;
            W.tc_ready = false;
            bool tc_ok = true;
            (void)tc_ok;
            {
                const std::vector<float> folded = pack_lower_frame(sd);
                ok &= upload(h, folded, W.frame) && upload(h, pack_lower_frame_mma(folded), W.frame_mma);
            }
            const std::string gp = "keyEncoder.gcn.";
            ok &= upload(h, pack_data_bn(sd, gp), W.data_bn);
            const int ch[4] = {3, 32, 64, 128};
            for (int i = 0; i < 3; ++i) {
                HostGcnLayer L = pack_gcn_layer(sd, gp, i, ch[i], ch[i + 1]);
                ok &= upload(h, L.ahat, W.gcn[i].ahat) && upload(h, L.gconv, W.gcn[i].gconv) && upload(h, L.tconv, W.gcn[i].tconv);
                W.gcn[i].cin = L.cin;
                W.gcn[i].cout = L.cout;
#ifndef MMEGO_EMUL
                if (tc_supported()) tc_ok = tc_ok && tc_pack_gemm(h, L.gconv, W.tc_gconv[i]) && tc_pack_gemm(h, L.tconv, W.tc_tconv[i]);
#endif
            }
            HostPackedGemm fcn = pack_linear(sd.get(gp + "fcn.weight", 64 * 128), sd.get(gp + "fcn.bias", 64), 64, {128});
            ok &= upload(h, fcn, W.fcn);
#ifndef MMEGO_EMUL
            if (tc_supported()) tc_ok = tc_ok && tc_pack_gemm(h, fcn, W.tc_fcn);
            W.tc_ready = tc_supported() && tc_ok;
#endif
            for (int l = 0; l < 3; ++l) {
                HostSmallLstm s = pack_small_lstm(sd, "fusion.rnn_pk.", l, l == 0 ? 192 : 128);
                ok &= upload(h, s.ih, W.lstm[l].ih) && upload(h, s.whh, W.lstm[l].whh);
                ok &= upload(h, pack_small_lstm_mma(sd, "fusion.rnn_pk.", l, l == 0 ? 192 : 128), W.lstm[l].mma);
                W.lstm[l].in = s.in;
            }
            ok &= upload(h, pack_linear(sd.get("fusion.fc0.weight", 128 * 173), sd.get("fusion.fc0.bias", 128), 128, {128, 45}), W.fc0);
            ok &= upload(h, pack_linear(sd.get("fusion.fc1.weight", 64 * 128), sd.get("fusion.fc1.bias", 64), 64, {128}), W.fc1);
            ok &= upload(h, pack_linear(sd.get("fusion.fc2.weight", 42 * 64), sd.get("fusion.fc2.bias", 42), 42, {64}), W.fc2);
            {
                const std::vector<float> hb = pack_head_mma({{sd.get("fusion.fc0.weight", 128 * 173), sd.get("fusion.fc0.bias", 128), 128, 173},
                                                             {sd.get("fusion.fc1.weight", 64 * 128), sd.get("fusion.fc1.bias", 64), 64, 128},
                                                             {sd.get("fusion.fc2.weight", 42 * 64), sd.get("fusion.fc2.bias", 42), 42, 64}});
                if (hb.size() != lower_head_mma_words()) return fail(h, MMEGO_ESHAPE, "set_weights: head layout mismatch");
                ok &= upload(h, hb, W.head_mma);
            }
            W.ready = ok;
        } else {
            return fail(h, MMEGO_EINVAL, "set_weights: unknown net %d", net);
        }
    } catch (const std::exception& ex) {
        return fail(h, MMEGO_ESHAPE, "set_weights: %s", ex.what());
    }
    if (!ok) return fail(h, MMEGO_ENOMEM, "set_weights: device allocation or upload failed");
    return MMEGO_OK;
}

size_t mmego_workspace_bytes(const mmego_handle* h, int stage, int B, int L, int N, int n_imu) {
    if (!h || B <= 0 || L <= 0) return 0;
    (void)N;
    Carver c(nullptr);
    const long long Bc = B < h->imu_chunk ? B : h->imu_chunk;
    if (stage == MMEGO_STAGE_IMU) {
        c.off = imu_ws_bytes(h, B, Bc, L, n_imu);
    } else if (stage == MMEGO_STAGE_UPPER) {
        UpperWs w;
        plan_upper(c, B, L, w);
    } else if (stage == MMEGO_STAGE_LOWER || stage == MMEGO_STAGE_GCN) {
        LowerWs w;
        plan_lower(c, B, L, w, use_gcn_tc(h));
    } else if (stage == MMEGO_STAGE_PIPELINE) {
        // R, t, upper_l, lower_l + the largest stage workspace (stages run back to back and reuse it)
        const size_t F = (size_t)B * L;
        c.f(F * 9); c.f(F * 3); c.f(F * 45); c.f(F * 24);
        size_t best = 0;
        best = std::max(best, imu_ws_bytes(h, B, Bc, L, n_imu));
        {
            Carver s(nullptr); UpperWs w; plan_upper(s, B, L, w); best = std::max(best, s.off);
        }
        {
            Carver s(nullptr); LowerWs w; plan_lower(s, B, L, w, use_gcn_tc(h)); best = std::max(best, s.off);
        }
        c.f(best / sizeof(float) + 64);
    } else {
        return 0;
    }
    return c.off + 256;
}

int mmego_imu_forward(mmego_handle* h, const float* imu, float* R, float* t, int B, int L, int n_imu, void* ws,
                      size_t ws_bytes, void* stream) {
    Entry entry(h);
    if (int rc = check_dims(h, B, L, 1)) return rc;
    if (!imu || !R || !t || !ws) return fail(h, MMEGO_EINVAL, "imu_forward: NULL argument");
    if (!h->imu.ready) return fail(h, MMEGO_ESTATE, "imu_forward: IMU_Net weights were never set");
    if (n_imu <= 0 || n_imu > 64) return fail(h, MMEGO_ESHAPE, "imu_forward: n_imu must be in 1..64 (got %d)", n_imu);
    if (ws_bytes < mmego_workspace_bytes(h, MMEGO_STAGE_IMU, B, L, 0, n_imu))
        return fail(h, MMEGO_ENOMEM, "imu_forward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long Bc = B < h->imu_chunk ? B : h->imu_chunk;
    Carver c(ws);
    if (use_resident(h, B, L)) {
        ImuResWs w;
        plan_imu_res(c, B, L, n_imu, w);
        if (int rc = imu_res_forward(h, imu, R, t, B, L, n_imu, w, st)) return rc;
        CUDA_TRY(h, cudaGetLastError());
        return MMEGO_OK;
    }
#ifndef MMEGO_EMUL
    if (h->imu_gemm != 0) {
        if (!h->imu.tc_ready) return fail(h, MMEGO_ESTATE, "imu_forward: tensor-core weights are not packed");
        ImuTcWs w;
        plan_imu_tc(c, B, Bc, L, n_imu, w);
        for (long long b0 = 0; b0 < B; b0 += Bc) {
            const long long nb = (B - b0) < Bc ? (B - b0) : Bc;
            if (int rc = imu_chunk_fast_tc(h, imu + (size_t)b0 * L * n_imu * kImuFeat, b0 * L, nb, L, n_imu, w, st)) return rc;
        }
        if (int rc = imu_slow_decode_tc(h, R, t, B, L, w, st)) return rc;
        CUDA_TRY(h, cudaGetLastError());
            return MMEGO_OK;
    }
#endif
#ifdef MMEGO_FFMA_GEN
    ImuWs w;
    plan_imu(c, Bc, L, n_imu, w);
    for (long long b0 = 0; b0 < B; b0 += Bc) {
        const long long nb = (B - b0) < Bc ? (B - b0) : Bc;
        imu_chunk_forward(h, imu + (size_t)b0 * L * n_imu * kImuFeat, R + (size_t)b0 * L * 9, t + (size_t)b0 * L * 3, nb, L,
                          n_imu, w, st);
    }
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
#else
    (void)st; (void)c; (void)Bc;
    return fail(h, MMEGO_ESTATE, "imu_forward: no kernel path for imu_gemm=%d in this build", h->imu_gemm);
#endif
}

int mmego_upper_forward(mmego_handle* h, float* x, const float* h0, const float* c0, const float* initial_body,
                        const float* R, const float* t, float* l, float* q, float* global_w, float* hn, float* cn,
                        int B, int L, int N, int body_index_mode, int b_offset, int B_global, void* ws,
                        size_t ws_bytes, void* stream) {
    Entry entry(h);
    if (int rc = check_dims(h, B, L, N)) return rc;
    if (!x || !initial_body || !R || !t || !l || !ws) return fail(h, MMEGO_EINVAL, "upper_forward: NULL argument");
    if (!h->upper.ready) return fail(h, MMEGO_ESTATE, "upper_forward: Upper_Net weights were never set");
    if (body_index_mode != MMEGO_BODY_REF && body_index_mode != MMEGO_BODY_PER_SNIPPET)
        return fail(h, MMEGO_EINVAL, "upper_forward: bad body_index_mode %d", body_index_mode);
    if (B_global < B + b_offset || b_offset < 0) return fail(h, MMEGO_EINVAL, "upper_forward: bad shard (b_offset %d, B %d, B_global %d)", b_offset, B, B_global);
    if (ws_bytes < mmego_workspace_bytes(h, MMEGO_STAGE_UPPER, B, L, N, 0))
        return fail(h, MMEGO_ENOMEM, "upper_forward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long F = (long long)B * L;
    Carver c(ws);
    UpperWs w;
    plan_upper(c, B, L, w);
    const UpperWeights& W = h->upper;
    {
        Prof p(h, "upper.point", st);
#ifdef MMEGO_FFMA_GEN
        if (!h->point_gemm)
            launch_upper_point(x, R, t, W.point.p, w.g, global_w, F, N, h->sm_count, st);
        else
#endif
            launch_upper_point_mma(x, R, t, W.point_mma.p, w.g, global_w, F, N, h->sm_count, h->point_stage, st);   // Upper_Net.py:379-381 (+gpointnet)
    }
    tap(h, "upper.g", w.g, (size_t)F * 64 * 4, st);
    const float* hs = run_small_lstm(h, W.lstm, w.g, 64, h0, c0, hn, cn, B, L, w.lstm, st);   // :339
    tap(h, "upper.lstm", hs, (size_t)F * 128 * 4, st);
    Prof p(h, "upper.head_decode", st);
#ifdef MMEGO_FFMA_GEN
    if (!h->head_gemm) {
        linear(h, W.fc1, hs, 128, w.h1, 128, F, 1, st);
        linear(h, W.fc2, w.h1, 128, w.o, 87, F, 0, st);
        tap(h, "upper.o", w.o, (size_t)F * 87 * 4, st);
        launch_upper_decode(w.o, initial_body, R, t, l, q, F, L, body_index_mode, (long long)b_offset * L, B_global, st);
    } else
#endif
    {
        // MLPHead + 6D -> rotations + forward kinematics + Transform2R in ONE kernel (:351-353, 355-387): the 87 head
        // outputs stay in shared memory (written out only for a debug tap)
        HeadTail tl{};
        tl.body = initial_body; tl.R = R; tl.t = t; tl.l = l; tl.q = q;
        tl.L = L; tl.mode = body_index_mode; tl.B_global = B_global; tl.row_offset = (long long)b_offset * L;
        const bool want_o = h->taps.count("upper.o") != 0;
        launch_upper_tail_mma(hs, W.head_mma.p, want_o ? w.o : nullptr, F, tl, h->sm_count, st);
        tap(h, "upper.o", w.o, (size_t)F * 87 * 4, st);
    }
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
}

}  // extern "C"

namespace {
// LowerNet.forward; with `assemble` the fused tail kernel also scatters upper_l / lower_l into pred [F,21,3] and ADDS the
// batch's error sums (target / pred / sums may each be null) -- Processor/Test/Demo_test.py:121-123, 64-69.
int lower_forward_impl(mmego_handle* h, const float* upper_l, float* x, const float* initial_body, const float* R,
                       const float* t, float* l, float* q, int B, int L, int N, int body_index_mode, int b_offset,
                       int B_global, void* ws, size_t ws_bytes, void* stream, int assemble, const float* target,
                       float* pred, double* sums, bool* assembled) {
    if (assembled) *assembled = false;
    if (int rc = check_dims(h, B, L, N)) return rc;
    if (!upper_l || !x || !initial_body || !R || !t || !l || !ws) return fail(h, MMEGO_EINVAL, "lower_forward: NULL argument");
    if (!h->lower.ready) return fail(h, MMEGO_ESTATE, "lower_forward: Lower_Net weights were never set");
    if (N < kLowerPts || N > lower_frame_max_points())
        return fail(h, MMEGO_ESHAPE, "lower_forward: N must be in %d..%d (got %d)", kLowerPts, lower_frame_max_points(), N);
    if (body_index_mode != MMEGO_BODY_REF && body_index_mode != MMEGO_BODY_PER_SNIPPET)
        return fail(h, MMEGO_EINVAL, "lower_forward: bad body_index_mode %d", body_index_mode);
    if (B_global < B + b_offset || b_offset < 0) return fail(h, MMEGO_EINVAL, "lower_forward: bad shard (b_offset %d, B %d, B_global %d)", b_offset, B, B_global);
    if (ws_bytes < mmego_workspace_bytes(h, MMEGO_STAGE_LOWER, B, L, N, 0))
        return fail(h, MMEGO_ENOMEM, "lower_forward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long F = (long long)B * L;
    Carver c(ws);
    LowerWs w;
    const bool gtc = use_gcn_tc(h);
    plan_lower(c, B, L, w, gtc);
    const LowerWeights& W = h->lower;
    if (gtc) {
#ifndef MMEGO_EMUL
        if (!W.tc_ready) return fail(h, MMEGO_ESTATE, "lower_forward: tensor-core ST-GCN weights are not packed");
        Prof p(h, "lower.gcn", st);
        tc_gcn_prep(upper_l, R, t, W.data_bn.p, w.uh, w.p_y0[0], w.p_y0[1], F, st);       // Lower_Net.py:229, GCN.py:339-344
        if (run_gcn_tc(h, B, L, w, st)) return fail(h, MMEGO_ECUDA, "lower_forward: tensor-core ST-GCN launch failed");
#endif
    } else {
#ifdef MMEGO_FFMA_GEN
        float* y0 = w.y[1];   // layer 0 writes y[0], so y[1] is free to hold the 3-channel input
        Prof p(h, "lower.gcn", st);
        launch_gcn_prep(upper_l, R, t, W.data_bn.p, w.uh, y0, F, st);                     // Lower_Net.py:229, GCN.py:339-344
        run_gcn(h, y0, B, L, w, st);
#else
        return fail(h, MMEGO_ESTATE, "lower_forward: no ST-GCN kernel path in this build");
#endif
    }
    tap(h, "lower.uh", w.uh, (size_t)F * 45 * 4, st);
    tap(h, "lower.K", w.kf, (size_t)F * kGcnV * 64 * 4, st);
    {
        Prof p(h, "lower.frame", st);
#ifdef MMEGO_FFMA_GEN
        if (!h->point_gemm)
            launch_lower_frame(x, R, t, w.kf, W.frame.p, w.ak, F, N, h->sm_count, st);
        else
#endif
            launch_lower_frame_mma(x, R, t, w.kf, W.frame_mma.p, w.ak, F, N, h->sm_count, st);   // :191-192, 216-227, 231, 104-116
    }
    tap(h, "lower.ak", w.ak, (size_t)F * 192 * 4, st);
    const float* hs = run_small_lstm(h, W.lstm, w.ak, 192, nullptr, nullptr, nullptr, nullptr, B, L, w.lstm, st);   // :117
    tap(h, "lower.lstm", hs, (size_t)F * 128 * 4, st);
    Prof p(h, "lower.head_decode", st);
#ifdef MMEGO_FFMA_GEN
    if (!h->head_gemm) {
        GemmArgs a = gemm_begin(W.fc0, w.f0, 128, F, 1);                                  // :119-121
        gemm_seg(a, W.fc0, hs, 128);
        gemm_seg(a, W.fc0, w.uh, 45);
        run_gemm(h, a, EPI_STORE, st);
        linear(h, W.fc1, w.f0, 128, w.f1, 64, F, 1, st);                                  // :122-123
        linear(h, W.fc2, w.f1, 64, w.o, 42, F, 0, st);                                    // :124
        tap(h, "lower.o", w.o, (size_t)F * 42 * 4, st);
        launch_lower_decode(w.o, initial_body, R, t, l, q, F, L, body_index_mode, (long long)b_offset * L, B_global, st);
    } else
#endif
    {
        // fc0-fc2 + 6D -> rotations + forward kinematics + Transform2R (+ 21-joint assembly + error sums) in ONE kernel
        // (:119-135, 235-238; Demo_test.py:121-123, 64-69)
        HeadTail tl{};
        tl.body = initial_body; tl.R = R; tl.t = t; tl.l = l; tl.q = q;
        tl.L = L; tl.mode = body_index_mode; tl.B_global = B_global; tl.row_offset = (long long)b_offset * L;
        tl.assemble = assemble; tl.upper_l = upper_l; tl.target = target; tl.pred = pred; tl.sums = sums;
        const bool want_o = h->taps.count("lower.o") != 0;
        launch_lower_tail_mma(hs, w.uh, W.head_mma.p, want_o ? w.o : nullptr, F, tl, h->sm_count, st);
        tap(h, "lower.o", w.o, (size_t)F * 42 * 4, st);
        if (assembled) *assembled = assemble != 0;
    }
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
}
}  // namespace

extern "C" {

int mmego_lower_forward(mmego_handle* h, const float* upper_l, float* x, const float* initial_body, const float* R,
                        const float* t, float* l, float* q, int B, int L, int N, int body_index_mode, int b_offset,
                        int B_global, void* ws, size_t ws_bytes, void* stream) {
    Entry entry(h);
    return lower_forward_impl(h, upper_l, x, initial_body, R, t, l, q, B, L, N, body_index_mode, b_offset, B_global, ws,
                              ws_bytes, stream, 0, nullptr, nullptr, nullptr, nullptr);
}

int mmego_gcn_extract_feature(mmego_handle* h, const float* x, float* out, int B, int T, void* ws, size_t ws_bytes,
                              void* stream) {
    Entry entry(h);
    if (int rc = check_dims(h, B, T, 1)) return rc;
    if (!x || !out || !ws) return fail(h, MMEGO_EINVAL, "gcn_extract_feature: NULL argument");
    if (!h->lower.ready) return fail(h, MMEGO_ESTATE, "gcn_extract_feature: Lower_Net weights were never set");
    if (ws_bytes < mmego_workspace_bytes(h, MMEGO_STAGE_GCN, B, T, 0, 0))
        return fail(h, MMEGO_ENOMEM, "gcn_extract_feature: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(ws);
    LowerWs w;
    const bool gtc = use_gcn_tc(h);
    plan_lower(c, B, T, w, gtc);
    if (gtc) {
#ifndef MMEGO_EMUL
        if (!h->lower.tc_ready) return fail(h, MMEGO_ESTATE, "gcn_extract_feature: tensor-core weights are not packed");
        tc_gcn_prep_raw(x, h->lower.data_bn.p, w.p_y0[0], w.p_y0[1], B, T, st);
        if (run_gcn_tc(h, B, T, w, st)) return fail(h, MMEGO_ECUDA, "gcn_extract_feature: tensor-core launch failed");
#endif
    } else {
#ifdef MMEGO_FFMA_GEN
        float* y0 = w.y[1];
        launch_gcn_prep_raw(x, h->lower.data_bn.p, y0, B, T, st);
        run_gcn(h, y0, B, T, w, st);
#else
        return fail(h, MMEGO_ESTATE, "gcn_extract_feature: no ST-GCN kernel path in this build");
#endif
    }
    CUDA_TRY(h, cudaMemcpyAsync(out, w.kf, (size_t)B * T * kGcnV * 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
}

int mmego_transform2h(mmego_handle* h, float* points, const float* R, const float* t, long long F, int n, int D,
                      void* stream) {
    Entry entry(h);
    if (!h) return MMEGO_EINVAL;
    if (!points || !R || !t || F < 0 || n <= 0 || D < 3) return fail(h, MMEGO_EINVAL, "transform2h: bad argument");
    launch_transform2h(points, R, t, F, n, D, static_cast<cudaStream_t>(stream));
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
}

int mmego_transform2r(mmego_handle* h, const float* points, const float* R, const float* t, float* out, long long F,
                      int n, void* stream) {
    Entry entry(h);
    if (!h) return MMEGO_EINVAL;
    if (!points || !R || !t || !out || F < 0 || n <= 0) return fail(h, MMEGO_EINVAL, "transform2r: bad argument");
    launch_transform2r(points, R, t, out, F, n, static_cast<cudaStream_t>(stream));
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
}

int mmego_assemble_metrics(mmego_handle* h, const float* upper_l, const float* lower_l, const float* target,
                           float* pred, double* sums, int B, int L, void* stream) {
    Entry entry(h);
    if (int rc = check_dims(h, B, L, 1)) return rc;
    if (!upper_l || !lower_l) return fail(h, MMEGO_EINVAL, "assemble_metrics: NULL argument");
    Prof p(h, "assemble_metrics", static_cast<cudaStream_t>(stream));
    launch_assemble_metrics(upper_l, lower_l, target, pred, sums, (long long)B * L, static_cast<cudaStream_t>(stream));
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
}

int mmego_pipeline_forward(mmego_handle* h, const float* imu, float* x, const float* initial_body, const float* target,
                           float* pred, double* sums, float* R_out, float* t_out, float* upper_out, float* lower_out,
                           int B, int L, int N, int n_imu, int body_index_mode, int b_offset, int B_global, void* ws,
                           size_t ws_bytes, void* stream) {
    Entry entry(h);
    if (int rc = check_dims(h, B, L, N)) return rc;
    if (!imu || !x || !initial_body || !ws) return fail(h, MMEGO_EINVAL, "pipeline_forward: NULL argument");
    if (ws_bytes < mmego_workspace_bytes(h, MMEGO_STAGE_PIPELINE, B, L, N, n_imu))
        return fail(h, MMEGO_ENOMEM, "pipeline_forward: workspace too small");
    const size_t F = (size_t)B * L;
    Carver c(ws);
    float* R = c.f(F * 9);
    float* t = c.f(F * 3);
    float* up = c.f(F * 45);
    float* lo = c.f(F * 24);
    float* sub = c.f(64);
    const size_t sub_bytes = ws_bytes - (size_t)(reinterpret_cast<char*>(sub) - static_cast<char*>(ws));
    if (R_out) R = R_out;
    if (t_out) t = t_out;
    if (upper_out) up = upper_out;
    if (lower_out) lo = lower_out;
    int rc = mmego_imu_forward(h, imu, R, t, B, L, n_imu, sub, sub_bytes, stream);
    if (rc) return rc;
    // h0 = c0 = 0 as built at Processor/Test/Demo_test.py:106-107
    rc = mmego_upper_forward(h, x, nullptr, nullptr, initial_body, R, t, up, nullptr, nullptr, nullptr, nullptr, B, L, N,
                             body_index_mode, b_offset, B_global, sub, sub_bytes, stream);
    if (rc) return rc;
    bool assembled = false;
    rc = lower_forward_impl(h, up, x, initial_body, R, t, lo, nullptr, B, L, N, body_index_mode, b_offset, B_global, sub,
                            sub_bytes, stream, 1, target, pred, sums, &assembled);
    if (rc) return rc;
    if (assembled) return MMEGO_OK;          // the 21-joint assembly and the error sums ran inside the fused lower tail
    return mmego_assemble_metrics(h, up, lo, target, pred, sums, B, L, stream);
}

int mmego_infer_host(mmego_handle* h, const float* imu_host, const float* data_host, const float* initial_body_host,
                     const float* target_host, float* pred_host, double* sums_host, int B, int L, int N, int n_imu,
                     int body_index_mode, int b_offset, int B_global) {
    Entry entry(h);
    if (int rc = check_dims(h, B, L, N)) return rc;
    if (!imu_host || !data_host || !initial_body_host) return fail(h, MMEGO_EINVAL, "infer_host: NULL argument");
    cudaSetDevice(h->device);
    if (!h->own_stream) CUDA_TRY(h, cudaStreamCreate(&h->own_stream));
    if (!h->h2d_stream) CUDA_TRY(h, cudaStreamCreate(&h->h2d_stream));
    if (!h->d2h_stream) CUDA_TRY(h, cudaStreamCreate(&h->d2h_stream));
    cudaStream_t st = h->own_stream;
    // The batch is cut into chunks of up to `host_chunk` snippets: chunk i+1 is copied to the device while chunk i
    // computes and chunk i-1's predictions travel back (three streams, two events per chunk).  The FIRST chunk is short
    // (its copy cannot overlap anything, so the pipeline should fill early), and every chunk size is picked so that
    // the work items of the dominant kernel (H=512 LSTM step: 16 per pair of 128-sequence tiles) fill whole rounds of
    // the sm_count/2 CTA pairs -- a 2048-snippet chunk would leave the last of its 35 rounds 40 % empty.
    const int Bc = (int)std::min<long long>(B, h->host_chunk);
    std::vector<std::pair<size_t, size_t>> chunks;      // (first snippet, count)
    {
        const long long pairs = std::max(1, h->sm_count / 2);
        auto efficiency = [&](long long nb) {
            const long long items = ((nb * L + 255) / 256) * 16;
            const long long rounds = (items + pairs - 1) / pairs;
            return (double)items / (double)(rounds * pairs) * ((double)nb * L / (double)(((nb * L + 255) / 256) * 256));
        };
        auto best = [&](long long lo, long long hi) {       // most efficient size in [lo, hi], larger wins ties
            long long arg = hi;
            double e = -1.0;
            for (long long nb = hi; nb >= lo && nb >= 1; --nb) {
                const double v = efficiency(nb);
                if (v > e + 1e-9) { e = v; arg = nb; }
            }
            return arg;
        };
        size_t b0 = 0;
        const size_t first = (size_t)Bc / 8;
        if (first >= 1 && (size_t)B > (size_t)Bc / 2 + first) {
            const size_t f = Bc >= 512 ? (size_t)best(Bc / 16, Bc / 4) : first;
            chunks.push_back({0, f});
            b0 = f;
        }
        while (b0 < (size_t)B) {
            size_t nb = std::min<size_t>(Bc, B - b0);
            if (B - b0 > (size_t)Bc && Bc >= 512) nb = (size_t)best(Bc - Bc / 8, Bc);
            chunks.push_back({b0, nb});
            b0 += nb;
        }
    }
    const int nchunks = (int)chunks.size();
    const size_t F = (size_t)B * L;
    const size_t ws_bytes = mmego_workspace_bytes(h, MMEGO_STAGE_PIPELINE, Bc, L, N, n_imu);
    Carver c(nullptr);
    // staging layout: imu, data, body, target, pred, sums(double), ws
    auto plan = [&](Carver& k, float*& imu, float*& data, float*& body, float*& tg, float*& pred, float*& sums, float*& ws) {
        imu = k.f(F * n_imu * kImuFeat);
        data = k.f(F * N * 6);
        body = k.f((size_t)B_global * 60);
        tg = k.f(F * 63);
        pred = k.f(F * 63);
        sums = k.f(2 * MMEGO_SUMS_LEN);
        ws = k.f(ws_bytes / sizeof(float) + 64);
    };
    float *imu, *data, *body, *tg, *pred, *sums, *ws;
    plan(c, imu, data, body, tg, pred, sums, ws);
    const size_t need = c.off + 256;
    if (need > h->stage_bytes) {
        if (h->stage_dev) cudaFree(h->stage_dev);
        h->stage_dev = nullptr;
        h->stage_bytes = 0;
        if (cudaMalloc(&h->stage_dev, need) != cudaSuccess) return fail(h, MMEGO_ENOMEM, "infer_host: cannot allocate %zu staging bytes", need);
        h->stage_bytes = need;
    }
    Carver k(h->stage_dev);
    plan(k, imu, data, body, tg, pred, sums, ws);
    while ((int)h->host_events.size() < 2 * nchunks) {
        cudaEvent_t e;
        CUDA_TRY(h, cudaEventCreate(&e));
        h->host_events.push_back(e);
    }
    const bool metrics = target_host && sums_host;
    CUDA_TRY(h, cudaMemcpyAsync(body, initial_body_host, (size_t)B_global * 60 * sizeof(float), cudaMemcpyHostToDevice, h->h2d_stream));
    if (metrics) CUDA_TRY(h, cudaMemsetAsync(sums, 0, MMEGO_SUMS_LEN * sizeof(double), h->h2d_stream));
    for (int i = 0; i < nchunks; ++i) {
        const size_t b0 = chunks[i].first, nb = chunks[i].second, f0 = b0 * L, nf = nb * L;
        CUDA_TRY(h, cudaMemcpyAsync(imu + f0 * n_imu * kImuFeat, imu_host + f0 * n_imu * kImuFeat,
                                    nf * n_imu * kImuFeat * sizeof(float), cudaMemcpyHostToDevice, h->h2d_stream));
        CUDA_TRY(h, cudaMemcpyAsync(data + f0 * N * 6, data_host + f0 * N * 6, nf * N * 6 * sizeof(float),
                                    cudaMemcpyHostToDevice, h->h2d_stream));
        if (metrics)
            CUDA_TRY(h, cudaMemcpyAsync(tg + f0 * 63, target_host + f0 * 63, nf * 63 * sizeof(float), cudaMemcpyHostToDevice,
                                        h->h2d_stream));
        CUDA_TRY(h, cudaEventRecord(h->host_events[2 * i], h->h2d_stream));
    }
    for (int i = 0; i < nchunks; ++i) {
        const size_t b0 = chunks[i].first, nb = chunks[i].second, f0 = b0 * L, nf = nb * L;
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->host_events[2 * i], 0));
        int rc = mmego_pipeline_forward(h, imu + f0 * n_imu * kImuFeat, data + f0 * N * 6, body, metrics ? tg + f0 * 63 : nullptr,
                                        pred + f0 * 63, metrics ? reinterpret_cast<double*>(sums) : nullptr, nullptr, nullptr,
                                        nullptr, nullptr, (int)nb, L, N, n_imu, body_index_mode, b_offset + (int)b0, B_global,
                                        ws, ws_bytes + 256, st);
        if (rc) return rc;
        CUDA_TRY(h, cudaEventRecord(h->host_events[2 * i + 1], st));
        if (pred_host) {
            CUDA_TRY(h, cudaStreamWaitEvent(h->d2h_stream, h->host_events[2 * i + 1], 0));
            CUDA_TRY(h, cudaMemcpyAsync(pred_host + f0 * 63, pred + f0 * 63, nf * 63 * sizeof(float), cudaMemcpyDeviceToHost,
                                        h->d2h_stream));
        }
    }
    if (metrics) CUDA_TRY(h, cudaMemcpyAsync(sums_host, sums, MMEGO_SUMS_LEN * sizeof(double), cudaMemcpyDeviceToHost, st));
    unsigned dev_err = 0;
    if (h->dev_error) CUDA_TRY(h, cudaMemcpyAsync(&dev_err, h->dev_error, sizeof(dev_err), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    CUDA_TRY(h, cudaStreamSynchronize(h->d2h_stream));
    if (dev_err)      // a kernel's bounded wait (item dependency / mbarrier) gave up: the outputs are not to be trusted
        return fail(h, MMEGO_ECUDA, "infer_host: a device-side wait timed out (flag %u); results are invalid", dev_err);
    return MMEGO_OK;
}

int mmego_build_snippets(mmego_handle* h, const mmego_raw_frames_t* raw, const long long* starts, const int* slot_src,
                         unsigned seed, float* data, float* imu, float* key, float* R, float* t, int B, int L, int N,
                         void* stream) {
    Entry entry(h);
    if (int rc = check_dims(h, B, L, N)) return rc;
    if (!raw || !starts || !data || !imu || !key || !R || !t) return fail(h, MMEGO_EINVAL, "build_snippets: NULL argument");
    if (!raw->points || !raw->pt_start || !raw->key || !raw->imu || !raw->R_btc || !raw->t_R0R || !raw->R_ref ||
        !raw->orientation_ref)
        return fail(h, MMEGO_EINVAL, "build_snippets: NULL array in mmego_raw_frames_t");
    if (raw->n_frames <= 0) return fail(h, MMEGO_EINVAL, "build_snippets: mmego_raw_frames_t.n_frames must be positive");
    RawFrames rf{raw->points, raw->pt_start, raw->key, raw->imu, raw->R_btc, raw->t_R0R, raw->R_ref, raw->orientation_ref,
                 raw->n_frames};
    Prof p(h, "build_snippets", static_cast<cudaStream_t>(stream));
    launch_snippet_build(rf, starts, slot_src, seed, data, imu, key, R, t, B, L, N, static_cast<cudaStream_t>(stream));
    CUDA_TRY(h, cudaGetLastError());
    return MMEGO_OK;
}

int mmego_debug_tap(mmego_handle* h, const char* name, void* dst, size_t bytes) {
    if (!h || !name) return MMEGO_EINVAL;
    if (!dst) { h->taps.erase(name); return MMEGO_OK; }
    h->taps[name] = {dst, bytes};
    return MMEGO_OK;
}

long long mmego_launch_count(const mmego_handle* h) { return h ? h->launches : 0; }

int mmego_profile_begin(mmego_handle* h) {
    if (!h) return MMEGO_EINVAL;
    for (ProfSpan& sp : h->prof) {
        if (sp.e0) cudaEventDestroy(sp.e0);
        if (sp.e1) cudaEventDestroy(sp.e1);
    }
    h->prof.clear();
    h->prof_on = true;
    return MMEGO_OK;
}

int mmego_profile_read(mmego_handle* h, const char* name, double* total_ms, long long* launches, long long* spans) {
    if (!h || !name) return MMEGO_EINVAL;
    double ms = 0.0;
    long long n = 0, k = 0;
    for (ProfSpan& sp : h->prof) {
        if (sp.name != name || !sp.e1) continue;
        CUDA_TRY(h, cudaEventSynchronize(sp.e1));
        float v = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&v, sp.e0, sp.e1));
        ms += v;
        n += sp.launches;
        ++k;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    if (spans) *spans = k;
    return MMEGO_OK;
}

int mmego_debug_stats(mmego_handle* h, unsigned long long* out8, int reset) {
    if (!h || !out8) return MMEGO_EINVAL;
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    if (h->dev_error) {       // out8[7]: set by a kernel whose bounded wait gave up (must stay 0)
        unsigned e = 0;
        CUDA_TRY(h, cudaDeviceSynchronize());
        unsigned long long raw[8];
        CUDA_TRY(h, cudaMemcpy(raw, h->dev_error, sizeof(raw), cudaMemcpyDeviceToHost));
        e = (unsigned)(raw[0] & 0xffffffffu);
        out8[7] = e;
        for (int i = 2; i < 7; ++i) out8[i - 2] = raw[i];      // cycle accounting of tconv_snip_kernel (test builds only)
        if (reset) CUDA_TRY(h, cudaMemset(h->dev_error, 0, 64));
    }
    if (!h->tc_stats) return MMEGO_OK;
    CUDA_TRY(h, cudaDeviceSynchronize());
    CUDA_TRY(h, cudaMemcpy(out8, h->tc_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(h, cudaMemset(h->tc_stats, 0, 8 * sizeof(unsigned long long)));
    return MMEGO_OK;
}

int mmego_profile_end(mmego_handle* h) {
    if (!h) return MMEGO_EINVAL;
    mmego_profile_begin(h);
    h->prof_on = false;
    return MMEGO_OK;
}

}  // extern "C"
