/* mmego_b200.h -- C ABI of libmmego_b200.so, the B200 (sm_100a) implementation of mmEgo's inference
 * forward pass (IMU_Net -> Upper_Net -> Lower_Net/ST-GCN -> joint decode + metrics).
 *
 * Every entry point names the reference interface it replaces (paths relative to the yenanjing/mmEgo
 * checkout).  Plain pointers and sizes only; no C++/torch types; no exceptions cross this boundary.
 *
 *  - All `float*` tensor arguments are DEVICE pointers to contiguous fp32 unless the name ends in `_host`.
 *  - Weights are passed as HOST pointers keyed by the checkpoint's state_dict names; the library folds
 *    BatchNorm, permutes LSTM gates and uploads its own packed copies (owned by the handle).
 *  - Calls are asynchronous on `stream` (a cudaStream_t passed as void*); nothing is retained past the call
 *    except packed weights.  A handle is not thread-safe; use one per GPU / per thread.
 *  - Return 0 on success, a negative MMEGO_E* code otherwise; text via mmego_last_error().
 *  - Numeric range of the default ("fp32-grade") mode: the tensor-core stages carry activations as two fp16 planes of
 *    2^s * value and SATURATE outside |value| <= 65000 / 2^s: 253 for IMU_Net's fc1 / LSTM / pooled features (s = 8; LSTM
 *    outputs are bounded by 1), 4062 for the ST-GCN activations (s = 4).  The shipped Upper/Lower checkpoints stay three
 *    orders of magnitude inside; a model whose activations leave the range is clipped, not turned into inf/NaN.
 */
#ifndef MMEGO_B200_H
#define MMEGO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMEGO_ABI_VERSION 2

enum {
    MMEGO_OK = 0,
    MMEGO_EINVAL = -1,  /* null pointer / bad argument / unknown option */
    MMEGO_ESHAPE = -2,  /* unsupported shape, missing or mis-sized weight tensor */
    MMEGO_EARCH = -3,   /* device is not sm_100 */
    MMEGO_ECUDA = -4,   /* CUDA runtime error */
    MMEGO_ENOMEM = -5,  /* workspace too small / allocation failure */
    MMEGO_ESTATE = -6   /* weights for this stage were never set */
};

enum { MMEGO_NET_IMU = 0, MMEGO_NET_UPPER = 1, MMEGO_NET_LOWER = 2 };
enum { MMEGO_STAGE_IMU = 0, MMEGO_STAGE_UPPER = 1, MMEGO_STAGE_LOWER = 2, MMEGO_STAGE_GCN = 3, MMEGO_STAGE_PIPELINE = 4 };

/* body_index_mode: which calibration skeleton flat frame r = b*L + l uses in forward kinematics.
 * 0 = reference-exact `initial_body[r % B]` (the `.repeat(L,1,1,1)` of Net/Upper_Net.py:134, Net/Lower_Net.py:26),
 * 1 = per-snippet `initial_body[r / L]`. */
enum { MMEGO_BODY_REF = 0, MMEGO_BODY_PER_SNIPPET = 1 };

/* Layout of the `sums` accumulator filled by mmego_assemble_metrics (all float64, ADDED to, never cleared):
 * [0..20] per-joint sum of ||pred-gt||, [21] upper-body sum (15 joints, UpperNet's own hips),
 * [22] lower-body sum (8 joints), [23..42] per-bone angle sum in degrees, [43] frame count,
 * [44] L1 sum |lower_l - gt| and [45] L1 sum over the 6 lower bone vectors (eval_loss / eval_loss_l, Demo_test.py:141-147). */
#define MMEGO_SUMS_LEN 46

typedef struct mmego_handle mmego_handle;

int mmego_abi_version(void);

/* Replaces model construction + `.to(device)` in Processor/Test/Demo_test.py:51-58.  Refuses non-sm_100 devices. */
int mmego_create(mmego_handle** out, int device);
int mmego_destroy(mmego_handle* h);
/* Text of the last error on this handle (or of the last failed mmego_create when h is NULL). */
const char* mmego_last_error(const mmego_handle* h);

/* Options: "imu_chunk"   (snippets per IMU_Net workspace chunk, default 2048),
 *          "imu_gemm"    (IMU_Net's H=512 LSTMs: 1 = tcgen05 fp16x3 split-precision tensor-core GEMM, fp32-grade, the
 *                         default; 2 = tcgen05 single-pass fp16, fastest, tolerance reported separately;
 *                         0 = fp32 FFMA GEMM),
 *          "gcn_gemm"    (ST-GCN GEMMs: 1 = tcgen05 fp16x3, default; 0 = fp32 FFMA),
 *          "point_gemm"  (radar point encoders + cross-attention: 1 = mma.sync fp16x3, default; 0 = fp32 FFMA),
 *          "small_lstm_gemm" (H=64 LSTMs: 1 = mma.sync fp16x3, default; 0 = fp32 FFMA),
 *          "gcn_kb_chunk" (gcn_gemm=1: K blocks of 64 accumulated in TMEM before draining into fp32 registers, default 4),
 *          "head_gemm"   (fully connected heads: 1 = one fused mma.sync kernel per head, default; 0 = fp32 FFMA GEMMs),
 *          "host_chunk"  (mmego_infer_host: snippets per stage of its H2D / compute / D2H pipeline, default 2048; the
 *                         first stage is an eighth of that so the un-overlappable first copy stays short),
 *          "imu_resident" (1, default: IMU_Net calls of at most "imu_res_max_seq" frames (B*L; default 120 = the measured
 *                         break-even, B <= 6 at L = 20) take the latency path -- gate weights resident in shared memory,
 *                         one persistent cooperative launch per bi-LSTM layer, 7 launches per call; 0 = always the
 *                         tcgen05 path),
 *          "imu_res_tc"  (latency path, 1 default: rnn_fast multiplies on mma.sync with
 *                         fp16 hi/lo split operands and fp32 accumulation, like every other GEMM of the library; 0 = exact
 *                         fp32 FMAs, 1.8x slower per step),
 *          "imu_res_direct" (latency path, 1 default: the mma.sync form reads its A fragments straight from L2, a contiguous
 *                         k range per warp and no barrier in the K loop; 0 = activations staged through a cp.async ring),
 *          "imu_res_xchg" (latency path, 1 default: layers with at most 4 sequences exchange h between the CTAs of a
 *                         direction as 64-bit words carrying a step tag -- no fence, arrival counter or poll; 0 = counter),
 *          "imu_res_pre" (latency path, 1 default: rnn_slow's layers take their input
 *                         projections for all timesteps up front; 0 = inside every timestep),
 *          "gcn_w_res"   (row-tiled ST-GCN GEMMs, 1 default: the weight matrix is loaded once per CTA and stays in shared
 *                         memory next to the activation ring when it fits),
 *          "gcn_snip"    (ST-GCN temporal convolutions, L <= 20: bit 0 = snippet-resident transposed kernel (one CTA per
 *                         snippet, window loaded once for the nine taps); bit 8+i = layer i stays on the row-tiled GEMM;
 *                         bit 4 / bit 12+i = a second accumulator drain per 64-channel block (all layers / layer i).
 *                         Default 513 = layers 0 and 2 snippet-resident, layer 1 row-tiled: the two kernels are equally
 *                         fast on the 64 -> 64 layer and the snippet kernel accumulates 108 MMAs per drain in TMEM, which
 *                         costs accuracy -- max lower-joint error over 200 snippets against the float64 oracle 6.4e-6 m,
 *                         9.6e-6 m with all three layers snippet-resident (value 1), 5.6e-6 m row-tiled (value 0, GCN
 *                         10 % slower): profiles/r02_gcn_accuracy_probe.json),
 *          "point_stage" (upper point encoder: 1 = radar clouds staged into shared memory by cp.async.bulk under the
 *                         previous frame's MMAs; default 0 -- measured 2-3 % slower than per-lane loads on a B200),
 *          "tc_kb_chunk" (imu_gemm=1: K blocks of 64 accumulated in TMEM before draining into fp32 registers, default 4;
 *                         longer chunks are NOT faster -- the MMA work is the limit -- and ~1.5x noisier at 8,
 *                         profiles/r01_lstm_chunk_accuracy.txt),
 *          "tc_cta_pair" (H=512 LSTM kernel on CTA pairs, cta_group::2 M=256 tiles, default 1),
 *          "tc_persist"  (H=512 LSTM kernel: timesteps 1..T-1 of a layer run as ONE launch whose work items wait for the
 *                         h_{t-1} they read; bit 0 = rnn_slow (default on: -1.1 ms per 4096 snippets, its 256 items per
 *                         step no longer pay 4 rounds for 3.46), bit 1 = rnn_fast (default off: that kernel is power
 *                         bound and measured slower without its idle tails); results are bit-identical either way),
 *          "tc_pdl"      (H=512 LSTM timestep launches use programmatic dependent launch: the prologue of step t+1
 *                         overlaps the tail of step t, default 1),
 *          "tc_lo_drop"  (imu_gemm=1: low mantissa bits rounded away in the residual (lo) fp16 planes of the H=512 LSTM
 *                         operands, 0..6, default 4 = 7 significant bits kept: the tensor core draws less power on
 *                         shorter operands and the kernel is power-limited; results are unchanged within the fp32-grade
 *                         tolerance up to 4.  The packed weight planes are rounded in place, so once IMU_Net is loaded
 *                         the value can only grow until the weights are loaded again). */
int mmego_set_option(mmego_handle* h, const char* key, long long value);

/* Replaces IMUNet.load / UpperNet.load / LowerNet.load (Net/IMU_Net.py:106-114, Net/Upper_Net.py:400-404,
 * Net/Lower_Net.py:251-258): `names[i]` is a state_dict key, `ptrs_host[i]` its fp32 data on the HOST,
 * `numels[i]` its element count.  Non-float entries (num_batches_tracked) and unused tensors may be omitted. */
int mmego_set_weights(mmego_handle* h, int net, const char* const* names, const float* const* ptrs_host,
                      const long long* numels, int n);

/* Bytes of device workspace the stage needs for these shapes (0 on error). */
size_t mmego_workspace_bytes(const mmego_handle* h, int stage, int B, int L, int N, int n_imu);

/* IMUNet.forward (Net/IMU_Net.py:67-94): imu [B,L,n_imu,15] -> R [B,L,3,3], t [B,L,3]. */
int mmego_imu_forward(mmego_handle* h, const float* imu, float* R, float* t, int B, int L, int n_imu, void* ws,
                      size_t ws_bytes, void* stream);

/* UpperNet.forward (Net/Upper_Net.py:374-388).  x [B,L,N,6] is IN-OUT: xyz is overwritten with R(xyz - t) exactly
 * as the reference's in-place Transform2H does.  h0,c0,hn,cn [6,B,64]; initial_body [B_global,20,3];
 * outputs l [B,L,15,3], q [B,L,14,3,3], global_w [B*L,N,1].  q/global_w/hn/cn may be NULL.
 * b_offset/B_global describe this shard's place in the unsharded batch (for MMEGO_BODY_REF). */
int mmego_upper_forward(mmego_handle* h, float* x, const float* h0, const float* c0, const float* initial_body,
                        const float* R, const float* t, float* l, float* q, float* global_w, float* hn, float* cn,
                        int B, int L, int N, int body_index_mode, int b_offset, int B_global, void* ws,
                        size_t ws_bytes, void* stream);

/* LowerNet.forward (Net/Lower_Net.py:177-239).  x [B,L,N,6] IN-OUT (second in-place Transform2H);
 * upper_l [B,L,15,3] is read only; outputs l [B,L,8,3], q [B,L,6,3,3] (q may be NULL).
 * Top-64 tie rule: among equal keys the lowest slot index wins (torch.sort(stable=True)). */
int mmego_lower_forward(mmego_handle* h, const float* upper_l, float* x, const float* initial_body, const float* R,
                        const float* t, float* l, float* q, int B, int L, int N, int body_index_mode, int b_offset,
                        int B_global, void* ws, size_t ws_bytes, void* stream);

/* GCN.Model.extract_feature (Net/GCN.py:332-355) with the Lower_Net checkpoint's keyEncoder.gcn.* weights:
 * x [B,3,T,15,1] -> out [B,T,15,64] (raw reinterpretation of the [B,64,T,15] block, as the reference). */
int mmego_gcn_extract_feature(mmego_handle* h, const float* x, float* out, int B, int T, void* ws, size_t ws_bytes,
                              void* stream);

/* Util/Universal_Util/Utils.py:284-292: points [F,n,D] xyz <- R (xyz - t), in place; R [F,3,3], t [F,3]. */
int mmego_transform2h(mmego_handle* h, float* points, const float* R, const float* t, long long F, int n, int D,
                      void* stream);
/* Util/Universal_Util/Utils.py:274-281: out [F,n,3] = R^T points + t. */
int mmego_transform2r(mmego_handle* h, const float* points, const float* R, const float* t, float* out, long long F,
                      int n, void* stream);

/* Processor/Test/Demo_test.py:121-123 + 64-69,150-158: scatters upper_l/lower_l into pred [B,L,21,3] (pred may be
 * NULL) and, when target [B,L,21,3] and sums are non-NULL, ADDS this batch's error sums into sums[MMEGO_SUMS_LEN]. */
int mmego_assemble_metrics(mmego_handle* h, const float* upper_l, const float* lower_l, const float* target,
                           float* pred, double* sums, int B, int L, void* stream);

/* The whole chain of Demo_test.py:111-123 on device buffers: IMU -> Upper -> Lower -> assemble (+metrics when
 * target/sums given).  x is IN-OUT as above.  R_out/t_out/upper_out/lower_out may be NULL (workspace is used). */
int mmego_pipeline_forward(mmego_handle* h, const float* imu, float* x, const float* initial_body, const float* target,
                           float* pred, double* sums, float* R_out, float* t_out, float* upper_out, float* lower_out,
                           int B, int L, int N, int n_imu, int body_index_mode, int b_offset, int B_global, void* ws,
                           size_t ws_bytes, void* stream);

/* Same chain with HOST buffers (the call a reference-side binding makes per batch): copies imu/data/skeleton/target
 * to the device, runs the pipeline, copies pred [B,L,21,3] and sums back -- chunk-pipelined on three streams so that the
 * copies of chunk i+1 / i-1 overlap the compute of chunk i.  Device staging is owned by the handle and grows on demand.
 * data_host is NOT modified.  Synchronous on return.  Host buffers should be pinned for the copies to overlap. */
int mmego_infer_host(mmego_handle* h, const float* imu_host, const float* data_host, const float* initial_body_host,
                     const float* target_host, float* pred_host, double* sums_host, int B, int L, int N, int n_imu,
                     int body_index_mode, int b_offset, int B_global);

/* Snippet builder -- replaces the per-frame arithmetic of the reference's loader (Util/Universal_Util/Dataset_sample.py:
 * 153-231: range channel and channel reorder :203-208, padding / sub-sampling to N slots :210-223, IMU re-framing
 * :184-195, R_R0R :182) and its snippet windows (:233-260, as the `starts` list) on raw per-frame sensor data kept
 * resident on the device (the packed cache of scripts/pack_sample_data.py).  All pointers are DEVICE pointers. */
typedef struct {
    const float* points;           /* [P][5] x, y, z, intensity, velocity of every radar point, frame after frame */
    const long long* pt_start;     /* [F+1] first point of every frame */
    const double* key;             /* [F][21][3] selected Kinect joints */
    const double* imu;             /* [F][20][15] raw imu_save_l */
    const double* R_btc;           /* [F][3][3] */
    const double* t_R0R;           /* [F][3] */
    const double* R_ref;           /* [3][3] R_btc of the reference frame (the loader's st == 0 frame) */
    const double* orientation_ref; /* [3][3] orientation_imu_img of the reference frame */
    long long n_frames;            /* F; a snippet window that leaves [0, F) yields all-zero frames (never an out-of-bounds read) */
} mmego_raw_frames_t;
/* starts [B]: first source frame of every snippet.  slot_src [B*L*N] (or NULL): source point of every output slot, -1 =
 * empty; NULL = seeded random placement (the reference uses an unseeded RNG).  Outputs: data [B,L,N,6],
 * imu [B,L,20,15], key [B,L,21,3], R [B,L,3,3], t [B,L,3], all fp32 as Demo_test.py:95-109 casts them. */
int mmego_build_snippets(mmego_handle* h, const mmego_raw_frames_t* raw, const long long* starts, const int* slot_src,
                         unsigned seed, float* data, float* imu, float* key, float* R, float* t, int B, int L, int N,
                         void* stream);

/* Test hook: the next forward copies the named intermediate tensor into dst (device, `bytes` capacity). */
int mmego_debug_tap(mmego_handle* h, const char* name, void* dst, size_t bytes);
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
long long mmego_launch_count(const mmego_handle* h);


/* Measurement hooks (bench.py): between profile_begin and profile_end every named group of launches is bracketed by
 * CUDA events on the launching stream.  profile_read synchronises on the recorded events and returns the summed
 * duration, the number of kernel launches and the number of spans of `name`
 * ("imu.fc1", "imu.lstm_fast", "imu.lstm_slow", "imu.pool", "imu.decode", "upper.point", "small_lstm", "upper.head_decode",
 *  "lower.gcn", "lower.frame", "lower.head_decode", "assemble_metrics", "build_snippets"). */
int mmego_profile_begin(mmego_handle* h);
int mmego_profile_read(mmego_handle* h, const char* name, double* total_ms, long long* launches, long long* spans);
int mmego_profile_end(mmego_handle* h);
/* Cycle counters of the tensor-core LSTM kernel's instrumentation.  The instrumentation (option "tc_dbg") is compiled
 * only into test builds (-DMMEGO_DEBUG_SWITCHES); the product library rejects the option and returns zeros here. */
int mmego_debug_stats(mmego_handle* h, unsigned long long* out8, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MMEGO_B200_H */
