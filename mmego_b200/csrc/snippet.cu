// snippet.cu -- GPU snippet builder: everything the reference's loader COMPUTES per frame
// (Util/Universal_Util/Dataset_sample.py:153-231) from the raw sensor data of the packed cache written by
// scripts/pack_sample_data.py.  One CTA per output frame (b, l), source frame f = starts[b] + l:
//   * radar cloud [N,6]: x, y, z, range = ||xyz|| (float64, then rounded: :204-206), velocity, intensity (:208);
//     n < N points land in n distinct random slots (the rest stay zero), n >= N keeps N random points (:211-223).
//     The reference draws from numpy's unseeded global RNG; here slot s / point p gets the counter-based key
//     hash(seed, f, index) and ranks decide (ties by index), or the caller passes the placement (slot_src);
//   * IMU block [20,15]: rotation columns re-framed R_RI (O_ref^T R_NI) R_RI^T, gravity and sign fixes (:186-195),
//     in float64 without FMA contraction so the float32 results round like numpy's;
//   * R_R0R = R_ttb R_ref R_btc^T R_ttb^T (:182), head translation t_R0R, the 21 ground-truth joints.
// HBM-bound byte shuffling (about 3 KB in, 4.8 KB out per frame).
#include "internal.h"

namespace mmego {

namespace {

constexpr int ST = 128;

__device__ __forceinline__ unsigned slot_hash(unsigned seed, unsigned frame, unsigned i) {
    unsigned x = seed * 0x9E3779B1u + frame * 0x85EBCA77u + i * 0xC2B2AE3Du + 0x27D4EB2Fu;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 12;
    x *= 0x297A2D39u;
    x ^= x >> 15;
    return x;
}

// rank of item i among `count` hashed items (ascending key, ties by index)
__device__ __forceinline__ int hash_rank(unsigned seed, unsigned frame, int i, int count) {
    const unsigned k = slot_hash(seed, frame, (unsigned)i);
    int r = 0;
    for (int q = 0; q < count; ++q) {
        const unsigned kq = slot_hash(seed, frame, (unsigned)q);
        r += (kq < k || (kq == k && q < i)) ? 1 : 0;
    }
    return r;
}

__device__ __forceinline__ double dot3(double a0, double b0, double a1, double b1, double a2, double b2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}
// C = A B (row-major 3x3), products summed in k order like numpy's small-matrix matmul
__device__ __forceinline__ void mm3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = dot3(A[i * 3], B[j], A[i * 3 + 1], B[3 + j], A[i * 3 + 2], B[6 + j]);
}
__device__ __forceinline__ void tr3(const double* A, double* T) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}

__global__ void __launch_bounds__(ST) snippet_build_kernel(RawFrames raw, const long long* __restrict__ starts,
                                                           const int* __restrict__ slot_src, unsigned seed,
                                                           float* __restrict__ data, float* __restrict__ imu,
                                                           float* __restrict__ key, float* __restrict__ R,
                                                           float* __restrict__ t, int L, int N) {
    const long long r = blockIdx.x;                      // output frame b*L + l
    const long long f = starts[r / L] + (r % L);         // source frame
    const int tid = threadIdx.x;
    if (f < 0 || f >= raw.n_frames) {                    // window outside the raw table: defined output, no stray reads
        for (int s = tid; s < N * 6; s += ST) data[r * (long long)N * 6 + s] = 0.f;
        for (int s = tid; s < 20 * 15; s += ST) imu[r * 300 + s] = 0.f;
        for (int s = tid; s < 63; s += ST) key[r * 63 + s] = 0.f;
        if (tid < 9) R[r * 9 + tid] = 0.f;
        if (tid < 3) t[r * 3 + tid] = 0.f;
        return;
    }
    const long long p0 = raw.pt_start[f];
    const int n = (int)(raw.pt_start[f + 1] - p0);
    // ---- radar cloud ---------------------------------------------------------------------------------------
    float* out = data + r * (long long)N * 6;
    if (slot_src) {
        for (int s = tid; s < N; s += ST) {
            const int p = slot_src[r * N + s];
            float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (p >= 0 && p < n) {
                const float* src = raw.points + (p0 + p) * 5;
                const double x = src[0], y = src[1], z = src[2];
                v[0] = src[0]; v[1] = src[1]; v[2] = src[2];
                v[3] = (float)__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
                v[4] = src[4]; v[5] = src[3];
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) out[s * 6 + c] = v[c];
        }
    } else {
        for (int s = tid; s < N * 6; s += ST) out[s] = 0.f;
        __syncthreads();
        const int count = n < N ? N : n;                 // hashed items: slots (n < N) or points (n >= N)
        for (int i = tid; i < count; i += ST) {
            const int rank = hash_rank(seed, (unsigned)f, i, count);
            const int slot = n < N ? i : rank, p = n < N ? rank : i;
            if ((n < N && rank < n) || (n >= N && rank < N)) {
                const float* src = raw.points + (p0 + p) * 5;
                const double x = src[0], y = src[1], z = src[2];
                float* o = out + slot * 6;
                o[0] = src[0]; o[1] = src[1]; o[2] = src[2];
                o[3] = (float)__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
                o[4] = src[4]; o[5] = src[3];
            }
        }
    }
    // ---- IMU block: one sample per thread -----------------------------------------------------------------------
    if (tid < 20) {
        const double* s = raw.imu + (f * 20 + tid) * 15;
        double v[15];
#pragma unroll
        for (int c = 0; c < 15; ++c) v[c] = s[c];
        // R_NI[i][j] = v[3j + i]  (np.stack([v[0:3], v[3:6], v[6:9]], axis=2))
        double RN[9], OT[9], A[9], Bm[9], M[9], RIT[9];
        const double RI[9] = {0, 0, 1, 0, -1, 0, 1, 0, 0};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) RN[i * 3 + j] = v[3 * j + i];
        tr3(raw.orientation_ref, OT);
        mm3(OT, RN, A);
        mm3(RI, A, Bm);
        tr3(RI, RIT);
        mm3(Bm, RIT, M);
#pragma unroll
        for (int c = 0; c < 9; ++c) v[c] = M[c];         // rows of M
        v[11] = __dadd_rn(v[11], 9.8);
        v[10] = -v[10]; v[11] = -v[11];
        v[13] = -v[13]; v[14] = -v[14];
        float* o = imu + (r * 20 + tid) * 15;
#pragma unroll
        for (int c = 0; c < 15; ++c) o[c] = (float)v[c];
    }
    // ---- R_R0R, t_R0R, joints -----------------------------------------------------------------------------------
    if (tid == 32) {
        const double TTB[9] = {0, -1, 0, -1, 0, 0, 0, 0, -1};
        double A[9], BT[9], C[9], TT[9], D[9];
        mm3(TTB, raw.R_ref, A);
        tr3(raw.R_btc + f * 9, BT);
        mm3(A, BT, C);
        tr3(TTB, TT);
        mm3(C, TT, D);
#pragma unroll
        for (int c = 0; c < 9; ++c) R[r * 9 + c] = (float)D[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[r * 3 + c] = (float)raw.t_R0R[f * 3 + c];
    }
    if (tid >= 64 && tid < 64 + 63) key[r * 63 + tid - 64] = (float)raw.key[f * 63 + tid - 64];
}

}  // namespace

void launch_snippet_build(const RawFrames& raw, const long long* starts, const int* slot_src, unsigned seed, float* data,
                          float* imu, float* key, float* R, float* t, long long B, int L, int N, cudaStream_t st) {
    if (B <= 0 || L <= 0) return;
    MMEGO_LAUNCH(snippet_build_kernel, dim3((unsigned)(B * L)), dim3(ST), 0, st, raw, starts, slot_src, seed, data, imu,
                 key, R, t, L, N);
}

}  // namespace mmego
