// internal.h -- shared declarations of libmmego_b200 (host side + kernel launchers).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "cuda_compat.h"

struct mmego_handle;

// The first-generation fp32 FFMA kernels (exact-fp32 references for A/B numbers, and what the CPU emulator suite runs
// where the tensor-core paths cannot be emulated) are compiled only into the emulator build and into the test-only
// library variant (-DMMEGO_WITH_FFMA, tests/_variant_build/).  The product library holds ONE kernel set.
#if defined(MMEGO_EMUL) || defined(MMEGO_WITH_FFMA)
#define MMEGO_FFMA_GEN 1
#endif

namespace mmego {

extern thread_local long long t_launches;   // kernels launched by the calling thread (credited to a handle by api.cu: Entry)

// ------------------------------------------------------------------------------------------------
// model constants (Config/config.py:16-24 of the reference)
// ------------------------------------------------------------------------------------------------
constexpr int kJointsAll = 21, kJointsUpper = 15, kJointsLower = 8, kBones = 20;
constexpr int kLowerPts = 64;        // Config.lower_pc_no
constexpr int kImuFeat = 15, kImuH = 512;
constexpr int kSmallH = 64;          // hidden size of grnn / rnn_pk
constexpr int kGcnV = 15;

// ------------------------------------------------------------------------------------------------
// generic fused GEMM (gemm_ffma.cu):  C[M,N] = sum_seg A_seg[M,K_seg] * W[N, Kp]^T  (+ epilogue)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSeg = 10;
struct GemmSeg {
    const float* a;      // row-major activations
    long long lda;       // row stride (floats)
    int k;               // valid K of this segment
    int kpad;            // K padded to a multiple of 16 (weight layout)
    int shift;           // source row = row + shift (temporal taps of the ST-GCN)
    int period;          // rows per snippet for the shift validity test (0 = no test)
    int vec;             // 1 = float4 loads allowed (aligned base, lda%4==0, k%8==0)
};
enum GemmEpi { EPI_STORE = 0, EPI_LSTM = 1, EPI_F6 = 2 };
struct GemmArgs {
    GemmSeg seg[kMaxSeg];
    int nseg;
    const float* w;      // packed [N][ldw], ldw = sum kpad of ALL packed segments
    int ldw;
    int ktot;            // sum kpad of the segments used by this call (a prefix of the packed ones)
    const float* bias;   // [N] or [rowmod][N]
    float* c;
    long long ldc;
    int M, N;
    int relu;
    int rowmod;          // >0: bias row = row % rowmod
    // EPI_LSTM
    float* cstate;       // [M][ldcs], updated in place
    long long ldcs;
    int has_state;       // 0: c_prev = 0 (first step with zero init)
    // EPI_F6: rows = b*period + pos; element (row, col) -> c[b*N*period + col*period + pos]
    int f6_period;
};
struct GemmBatch {
    GemmArgs g[2];
};
// bn: tile width (32, 64, 128); nz: 1 or 2 problems (blockIdx.z)
void launch_gemm(const GemmBatch& b, int nz, int bn, int epi, cudaStream_t st);

void handle_free(mmego_handle* h, void* p);   // cudaFree + forget a device buffer owned by the handle (api.cu)

// per-device one-shot flag (kernel attributes such as the dynamic shared memory limit are per device)
inline bool first_use_on_device(bool* flags) {
    int d = 0;
    cudaGetDevice(&d);
    if (flags[d & 63]) return false;
    flags[d & 63] = true;
    return true;
}

// other kernel launchers (one per .cu file)
size_t upper_point_smem_bytes();
void launch_upper_point(float* x, const float* R, const float* t, const float* wblob, float* g, float* gw, long long F,
                        int N, int sm_count, cudaStream_t st);
void launch_lstm_small(const float* gx, const float* whh, const float* h0, const float* c0, float* y, float* hn,
                       float* cn, int S, int T, cudaStream_t st);
void launch_upper_point_mma(float* x, const float* R, const float* t, const float* wblob, float* g, float* gw,
                            long long F, int N, int sm_count, int stage_clouds, cudaStream_t st);
void launch_lstm_small_mma(const float* x, long long ldx, int In, const float* blob, float* gx, const float* h0,
                           const float* c0, float* y, float* hn, float* cn, int S, int T, int sm_count,
                           cudaStream_t st);
// arguments of the decode stage fused behind the fully connected heads (heads_mma.cu)
struct HeadTail {
    const float* body;       // initial_body [B_global,20,3]
    const float* R;          // [F,3,3]
    const float* t;          // [F,3]
    float* l;                // joints out: [F,15,3] (upper) / [F,8,3] (lower)
    float* q;                // rotations out or null: [F,14,3,3] / [F,6,3,3]
    int L, mode, B_global;   // body_index_mode and the shard's place in the global batch
    long long row_offset;
    // lower tail only: 21-joint assembly + error sums (Processor/Test/Demo_test.py:121-123, 64-69)
    int assemble;
    const float* upper_l;    // [F,15,3]
    const float* target;     // [F,21,3] or null
    float* pred;             // [F,21,3] or null
    double* sums;            // [MMEGO_SUMS_LEN] or null
};
void launch_upper_tail_mma(const float* x, const float* blob, float* o, long long F, const HeadTail& tail, int sm_count,
                           cudaStream_t st);
void launch_lower_tail_mma(const float* hs, const float* uh, const float* blob, float* o, long long F, const HeadTail& tail,
                           int sm_count, cudaStream_t st);
size_t upper_head_mma_words();
size_t lower_head_mma_words();
size_t lower_frame_smem_bytes();
int lower_frame_max_points();
void launch_lower_frame(float* x, const float* R, const float* t, const float* kfeat, const float* wblob, float* ak,
                        long long F, int N, int sm_count, cudaStream_t st);
void launch_lower_frame_mma(float* x, const float* R, const float* t, const float* kfeat, const float* wblob,
                            float* ak, long long F, int N, int sm_count, cudaStream_t st);
void launch_gcn_prep(const float* upper, const float* R, const float* t, const float* bn, float* uh, float* y0,
                     long long F, cudaStream_t st);
void launch_gcn_prep_raw(const float* x, const float* bn, float* y0, int B, int T, cudaStream_t st);
void launch_gcn_agg(const float* y, const float* ahat, float* ya, long long F, int C, int sm_count, cudaStream_t st);
void launch_upper_decode(const float* o, const float* body, const float* R, const float* t, float* l, float* q,
                         long long F, int L, int mode, long long row_offset, int B_global, cudaStream_t st);
void launch_lower_decode(const float* o, const float* body, const float* R, const float* t, float* l, float* q,
                         long long F, int L, int mode, long long row_offset, int B_global, cudaStream_t st);
void launch_assemble_metrics(const float* up, const float* lo, const float* tg, float* pred, double* sums, long long F,
                             cudaStream_t st);
void launch_imu_pool(const float* y, const float* attn, float* out, long long F, int n, cudaStream_t st);
void launch_imu_decode(const float* g, const float* fc2, float* R, float* t, long long F, cudaStream_t st);
void launch_transform2h(float* pts, const float* R, const float* t, long long F, int n, int D, cudaStream_t st);
void launch_transform2r(const float* pts, const float* R, const float* t, float* out, long long F, int n,
                        cudaStream_t st);

// small-batch latency path of IMU_Net (lstm_resident.cu): fp32, gate weights resident in shared memory across timesteps
constexpr int kResMaxSeq = 120;      // B*L up to which the resident path is taken (B <= 6 at L = 20; measured break-even with the
                                     // tcgen05 path: B = 7, profiles/r02_latency_break_even_final.txt); run-time: option imu_res_max_seq
struct StateDict;
void pack_resident_layer(const StateDict& sd, const std::string& prefix, int layer, int In, std::vector<float>& w,
                         std::vector<float>& bias);
bool resident_supported(int sm_count);
size_t resident_gx_floats(int S, int T);
void pack_resident_layer_tc(const StateDict& sd, const std::string& prefix, int layer, int In, std::vector<float>& w,
                            std::vector<float>& scale);
size_t resident_xchg_words(int S);
int launch_lstm_resident(const float* x, int In, float* y, const float* w, const float* wscale, const float* bias,
                         float* cstate, unsigned* flags, float* gxs, unsigned long long* xchg, int direct, int S, int T,
                         cudaStream_t st);
void launch_res_fc1(const float* imu, const float* w, float* u, long long rows, cudaStream_t st);

// snippet builder (snippet.cu): device views of the packed raw cache (scripts/pack_sample_data.py)
struct RawFrames {
    const float* points;            // [P][5]  x, y, z, intensity, velocity
    const long long* pt_start;      // [F+1]
    const double* key;              // [F][21][3]
    const double* imu;              // [F][20][15]
    const double* R_btc;            // [F][3][3]
    const double* t_R0R;            // [F][3]
    const double* R_ref;            // [3][3]
    const double* orientation_ref;  // [3][3]
    long long n_frames;             // F
};
void launch_snippet_build(const RawFrames& raw, const long long* starts, const int* slot_src, unsigned seed, float* data,
                          float* imu, float* key, float* R, float* t, long long B, int L, int N, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// packed weights
// ------------------------------------------------------------------------------------------------
struct TcLstmLayer;
struct PackedGemm;
struct DevBuf {
    float* p = nullptr;
    size_t n = 0;
};
struct PackedGemm {     // W [N][ldw] + bias
    DevBuf w, bias;
    int N = 0, ldw = 0, nseg = 0;
    int k[kMaxSeg] = {0}, kpad[kMaxSeg] = {0};
};
struct PackedBigLstmLayer {   // one layer, both directions, H=512, gate-interleaved rows (tile of 32 units)
    PackedGemm dir[2];
};
struct PackedSmallLstmLayer {  // H=64
    PackedGemm ih;             // [512][In] both dirs, column order = recurrent-kernel thread order; bias = b_ih+b_hh
    DevBuf whh;                // [2][256][64]  row = thread order
    DevBuf mma;                // both GEMMs as mma.sync fragments (pack_small_lstm_mma)
    int in = 0;
};
// H=512 layer packed for the tcgen05 path (lstm_tc.cu): fp16 hi/lo planes [2 dirs * 2048 rows][In + 512], scaled by 2^e,
// rows gate-interleaved per 256-column tile (64 hidden units)
struct TcLstmVariant {
    alignas(64) unsigned char map_hi[128];
    alignas(64) unsigned char map_lo[128];
    alignas(64) unsigned char map_hi2[128];    // same tensors, box of half the rows: one CTA's half of a CTA pair's tile
    alignas(64) unsigned char map_lo2[128];
    void* whi = nullptr;
    void* wlo = nullptr;
    float* bias = nullptr;     // [2][2048] packed row order
};
struct TcLstmLayer {
    TcLstmVariant v[1];
    int in_features = 0, K = 0;
    float out_scale = 1.f;     // 2^-e
};

// a GEMM weight packed for gemm_tc.cu: fp16 hi/lo planes [N][K64] of 2^e W, every K segment padded to 64
struct TcGemmW {
    alignas(64) unsigned char map_hi[128];
    alignas(64) unsigned char map_lo[128];
    void* whi = nullptr;
    void* wlo = nullptr;
    float* bias = nullptr;
    int N = 0, K64 = 0;
    float out_scale = 1.f;
};

struct ImuWeights {
    bool ready = false;
    bool tc_ready = false;
    TcLstmLayer tc_fast[2], tc_slow[2];
    PackedGemm fc1;
    DevBuf fc1_mma;    // fc1 as mma.sync fragments (pack_imu_fc1_mma)
    PackedBigLstmLayer fast[2], slow[2];
    DevBuf attn;       // [1024] + [1] bias at the end
    DevBuf fc2;        // [9][1024] + [9]
    // latency path (lstm_resident.cu): fp32 k-major slices per (direction, 8-unit group); order fast l0, l1, slow l0, l1
    bool res_ready = false;
    DevBuf res_w[4], res_b[4];
    DevBuf res_wtc[2], res_stc[2];   // rnn_fast layers as mma.sync fragments + per-slice scales (tensor-core form of the latency path)
    DevBuf res_fc1;    // [512][15] + [512]
};
struct UpperWeights {
    bool ready = false;
    DevBuf point;      // folded per-point MLP blob (point_layout.h)
    DevBuf point_mma;  // the same network as mma.sync fragments (UpperMmaLayout)
    PackedSmallLstmLayer lstm[3];
    PackedGemm fc1, fc2;
    DevBuf head_mma;   // mlpHead.fc1/fc2 as mma.sync fragments (pack_head_mma)
};
struct GcnLayerWeights {
    DevBuf ahat;       // [2][15][15]
    PackedGemm gconv;  // [C'][2C] (+ rowmod bias [15][C'])
    PackedGemm tconv;  // [C'][9*C' + C]
    int cin = 0, cout = 0;
};
struct LowerWeights {
    bool ready = false;
    bool tc_ready = false;
    TcGemmW tc_gconv[3], tc_tconv[3], tc_fcn;
    DevBuf frame;      // folded per-point MLP + to_q/to_k/to_v blob (point_layout.h)
    DevBuf frame_mma;  // the same as mma.sync fragments (LowerMmaLayout)
    DevBuf data_bn;    // [45] scale, [45] offset
    GcnLayerWeights gcn[3];
    PackedGemm fcn;
    PackedSmallLstmLayer lstm[3];
    PackedGemm fc0, fc1, fc2;
    DevBuf head_mma;   // fusion.fc0/fc1/fc2 as mma.sync fragments
};

// host-side packing (pack.cpp; pure C++, unit-tested on CPU)
using HostSD = std::map<std::string, std::pair<const float*, long long>>;
struct HostPackedGemm {
    std::vector<float> w, bias;
    int N = 0, ldw = 0, nseg = 0;
    int k[kMaxSeg] = {0}, kpad[kMaxSeg] = {0};
};
inline int pad16(int k) { return (k + 15) / 16 * 16; }

#ifndef MMEGO_EMUL
// tcgen05 path (lstm_tc.cu; nvcc only)
struct StateDict;
bool tc_supported();
bool tc_pack_layer(mmego_handle* h, const StateDict& sd, const std::string& prefix, int layer, int In, TcLstmLayer& out);
int tc_lstm_layer(mmego_handle* h, const TcLstmLayer& lw, const void* xhi, const void* xlo, void* yhi, void* ylo,
                  float* cstate, long long S, long long Spad, int T, int npass, bool persist_steps, cudaStream_t st);
void tc_imu_fc1(const float* imu, const float* fc1_mma, void* uhi, void* ulo, long long rows, int sm_count, int lo_drop,
                cudaStream_t st);
void tc_imu_pool(const void* yhi, const void* ylo, const float* attn, void* shi, void* slo, long long F, int n,
                 int lo_drop, cudaStream_t st);
void tc_round_lo_weights(const TcLstmLayer& lw, int drop, cudaStream_t st);
void tc_imu_decode(const void* ghi, const void* glo, const float* fc2, float* R, float* t, long long F, cudaStream_t st);
void tc_unsplit(const void* hi, const void* lo, float* out, long long n, cudaStream_t st);
// ST-GCN on tensor cores (gemm_tc.cu)
bool tc_pack_gemm(mmego_handle* h, const HostPackedGemm& g, TcGemmW& out);
int tc_gcn_gemm(mmego_handle* h, const TcGemmW& w, const void* a0hi, const void* a0lo, int c0, int n0, const void* a1hi,
                const void* a1lo, int c1, int rowmod, int relu, void* outhi, void* outlo, float* out_f6, int B, int RP,
                cudaStream_t st);
bool tc_gcn_tconv_snip_supported(int L);
int tc_gcn_tconv_snip(mmego_handle* h, const TcGemmW& w, const void* uhi, const void* ulo, int cu, const void* yhi,
                      const void* ylo, int cy, int cy_real, void* outhi, void* outlo, int B, int L, int subdrain, cudaStream_t st);
void tc_gcn_prep(const float* upper, const float* R, const float* t, const float* bn, float* uh, void* yhi, void* ylo,
                 long long F, cudaStream_t st);
void tc_gcn_prep_raw(const float* x, const float* bn, void* yhi, void* ylo, int B, int T, cudaStream_t st);
void tc_gcn_agg(const void* yhi, const void* ylo, const float* ahat, void* ohi, void* olo, long long F, int C, int CS,
                int OS, int sm_count, cudaStream_t st);
#endif

}  // namespace mmego

struct ProfSpan {
    std::string name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    long long launches = 0;   // kernels launched by calls on THIS handle
    int entry_depth = 0;
};

struct mmego_handle {
    bool prof_on = false;
    std::vector<ProfSpan> prof;
    int device = 0;
    int sm_count = 148;
    std::string err;
    long long launches = 0;   // kernels launched by calls on THIS handle
    int entry_depth = 0;
    long long imu_chunk = 2048;
    int imu_gemm = -1;        // -1: pick at first use (1 when the tcgen05 path is available, else 0)
    void* tc_stats = nullptr; // device counters of the dbg & 4 instrumentation
    int tc_dbg = 0;           // experiment switches of lstm_tc.cu (never set in product use)
    int tc_cta_pair = 1;      // H=512 LSTM kernel: 1 = CTA pairs (cta_group::2, M = 256)
    int tc_lo_drop = 4;       // H=512 LSTM operands: low mantissa bits rounded away in the residual (lo) planes, 0..6 (power, lstm_tc.cu)
    int tc_lo_drop_w = 0;     // ... already applied to the packed weight planes (one-way)
    int tc_persist = 1;       // H=512 LSTM kernel: timesteps 1..T-1 of a layer as ONE launch with item-level dependencies; bit 0 = rnn_slow, bit 1 = rnn_fast (0 = one launch per step)
    unsigned* tc_sync = nullptr;   // its arrival counters
    size_t tc_sync_words = 0;
    int tc_pdl = 1;           // H=512 LSTM kernel: 1 = programmatic dependent launch (step t+1's prologue overlaps step t's tail)
    int gcn_gemm = 0;         // ST-GCN GEMMs: 0 = fp32 FFMA, 1 = tcgen05 fp16x3 (default when available)
    int point_gemm = 1;       // point encoders + cross attention: 0 = fp32 FFMA, 1 = mma.sync fp16x3 (default)
    int gcn_kb_chunk = 4;     // ST-GCN tcgen05 GEMMs: K blocks of 64 accumulated in TMEM before draining into fp32 registers (0 = all)
    int head_gemm = 1;        // fully connected heads: 0 = fp32 FFMA GEMMs, 1 = one fused mma.sync kernel per head (default)
    int small_lstm_gemm = 1;  // H=64 LSTMs: 0 = fp32 FFMA, 1 = mma.sync fp16x3 (default)
    int tc_kb_chunk0 = 8;     // ... of the first two chunks of every tile
    int tc_kb_chunk = 4;      // fp16x3 mode: K blocks (of 64) per TMEM partial accumulation (see lstm_tc.cu)
    int point_stage = 0;      // upper point encoder: 1 = radar clouds staged into shared memory by TMA bulk copies one frame ahead
    int gcn_w_res = 1;        // row-tiled ST-GCN GEMMs: weights resident in shared memory when they fit next to >= 3 activation stages
    int gcn_snip = 1 | (256 << 1);   // ST-GCN temporal convs: bit 0 = snippet-resident transposed kernel (L <= 20), bit 8+i = layer i stays on
                              // the row-tiled GEMM (default: the middle layer -- same speed there, and the snippet kernel's long
                              // TMEM accumulation chains cost accuracy), bit 4 / 12+i = second drain group (all layers / layer i)
    unsigned* dev_error = nullptr;   // device word set by a kernel whose bounded wait gave up (mmego_debug_stats out8[7])
    int imu_res_max_seq = mmego::kResMaxSeq;   // B*L up to which IMU_Net takes the latency path
    int imu_res_direct = 1;   // latency path, tensor-core form: A fragments straight from L2, no staging ring and no barrier in the K loop
    int imu_res_xchg = 1;     // latency path: h travels between the CTAs of a direction as tagged 64-bit words (no fence / arrival counter / poll)
    int imu_res_tc = 1;       // latency path: rnn_fast on mma.sync (fp16 hi/lo split, fp32 accumulate); 0 = exact fp32 FMAs
    int imu_res_pre = 1;      // latency path: rnn_slow's input projections of all timesteps up front
    int imu_resident = 1;     // small batches (B*L <= kResMaxSeq): persistent fp32 LSTM with weights resident in shared memory
    mmego::ImuWeights imu;
    mmego::UpperWeights upper;
    mmego::LowerWeights lower;
    std::map<std::string, std::pair<void*, size_t>> taps;
    std::vector<void*> owned;      // device allocations freed at destroy
    // host-API staging
    void* stage_dev = nullptr;
    size_t stage_bytes = 0;
    cudaStream_t own_stream = nullptr, h2d_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> host_events;
    long long host_chunk = 2048;   // snippets per H2D/compute/D2H pipeline stage of mmego_infer_host (first stage: 1/8)
};
