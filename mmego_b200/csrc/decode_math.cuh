// decode_math.cuh -- per-frame decode arithmetic shared by the standalone decode kernels (decode.cu) and the fused
// head + decode (+ assembly + metrics) kernels (heads_mma.cu): 6D -> rotation, forward kinematics along the reference's
// bone order, Transform2R, 21-joint assembly and the error terms of Processor/Test/Demo_test.py.
#pragma once
#include "cuda_compat.h"

namespace mmego {
namespace dec {

// skeleton tables (Config/config.py:37-55 of the reference): bone i = [parent, child] of skeleton_all
__constant__ int kSkelParent[20] = {20, 3, 2, 2, 2, 4, 5, 6, 8, 9, 10, 1, 0, 0, 12, 13, 14, 16, 17, 18};
__constant__ int kSkelChild[20] = {3, 2, 1, 4, 8, 5, 6, 7, 9, 10, 11, 0, 12, 16, 13, 14, 15, 17, 18, 19};

constexpr int kSumsLen = 46;   // == MMEGO_SUMS_LEN of the public header

// upper-local index of a 21-joint id: upper_joint_map = [0..12, 16, 20]
__host__ __device__ constexpr int upper_idx(int j) { return j <= 12 ? j : (j == 16 ? 13 : 14); }
// lower-local index: lower_joint_map = [12..19]
__host__ __device__ constexpr int lower_idx(int j) { return j - 12; }
// rotation slot of a lower child joint: [13,14,15,17,18,19].index(c)
__host__ __device__ constexpr int lower_rot_idx(int c) { return c <= 15 ? c - 13 : c - 14; }

// Gram-Schmidt 6D -> rotation, columns (x, y, z); m is row-major 3x3 (Net/Upper_Net.py:355-364, Net/IMU_Net.py:7-47).
__device__ __forceinline__ void ortho6d(const float* a6, float eps, float* m) {
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

// Transform2R (Util/Universal_Util/Utils.py:274-281) of n joints in place: J <- R^T J + t;  rt = [R row-major (9) | t (3)]
__device__ __forceinline__ void to_reference_frame(float* J, int n, const float* rt) {
    for (int j = 0; j < n; ++j) {
        float* ld = J + j * 3;
        const float jx = ld[0], jy = ld[1], jz = ld[2];
        ld[0] = rt[0] * jx + rt[3] * jy + rt[6] * jz + rt[9];
        ld[1] = rt[1] * jx + rt[4] * jy + rt[7] * jz + rt[10];
        ld[2] = rt[2] * jx + rt[5] * jy + rt[8] * jz + rt[11];
    }
}

// Upper_Net tail of ONE frame: in[87] (14 x 6D | head) -> J[45] (15 joints, reference frame), q (14 x 3x3, stride qs
// between matrices; may be null).  bd = this frame's 20 bone vectors (ForKinematics, Net/Upper_Net.py:122-144).
__device__ __forceinline__ void upper_frame_decode(const float* in, const float* bd, const float* rt, float* J, float* q) {
    J[42] = in[84]; J[43] = in[85]; J[44] = in[86];
    for (int i = 0; i < 14; ++i) {
        const int ci = upper_idx(kSkelChild[i]), pi = upper_idx(kSkelParent[i]);
        float m[9];
        ortho6d(in + ci * 6, 1e-12f, m);
        if (q) {
#pragma unroll
            for (int k = 0; k < 9; ++k) q[ci * 9 + k] = m[k];
        }
        const float bx = bd[i * 3], by = bd[i * 3 + 1], bz = bd[i * 3 + 2];
        J[ci * 3] = J[pi * 3] + (m[0] * bx + m[1] * by + m[2] * bz);
        J[ci * 3 + 1] = J[pi * 3 + 1] + (m[3] * bx + m[4] * by + m[5] * bz);
        J[ci * 3 + 2] = J[pi * 3 + 2] + (m[6] * bx + m[7] * by + m[8] * bz);
    }
    to_reference_frame(J, 15, rt);
}

// Lower_Net tail of ONE frame: in[42] (6 x 6D | hip_l | hip_r) -> J[24] (8 joints), q (6 x 3x3; may be null)
// (FusionModule tail Net/Lower_Net.py:126-135, ForKinematics :12-37).
__device__ __forceinline__ void lower_frame_decode(const float* in, const float* bd, const float* rt, float* J, float* q) {
    J[0] = in[36]; J[1] = in[37]; J[2] = in[38];
    J[12] = in[39]; J[13] = in[40]; J[14] = in[41];
    for (int i = 0; i < 6; ++i) {
        const int c = kSkelChild[14 + i], p = kSkelParent[14 + i];
        const int ci = lower_idx(c), pi = lower_idx(p), qi = lower_rot_idx(c);
        float m[9];
        ortho6d(in + qi * 6, 1e-12f, m);
        if (q) {
#pragma unroll
            for (int k = 0; k < 9; ++k) q[qi * 9 + k] = m[k];
        }
        const float bx = bd[(14 + i) * 3], by = bd[(14 + i) * 3 + 1], bz = bd[(14 + i) * 3 + 2];
        J[ci * 3] = J[pi * 3] + (m[0] * bx + m[1] * by + m[2] * bz);
        J[ci * 3 + 1] = J[pi * 3 + 1] + (m[3] * bx + m[4] * by + m[5] * bz);
        J[ci * 3 + 2] = J[pi * 3 + 2] + (m[6] * bx + m[7] * by + m[8] * bz);
    }
    to_reference_frame(J, 8, rt);
}

// pred[:, :, upper_joint_map] = upper ; pred[:, :, lower_joint_map] = lower -- lower wins on joints 12 and 16
// (Processor/Test/Demo_test.py:121-123).  u[45], l[24] -> p[63]
__device__ __forceinline__ void assemble_frame(const float* u, const float* l, float* p) {
#pragma unroll
    for (int j = 0; j < 21; ++j) {
        const float* src = (j >= 12 && j <= 19) ? (l + (j - 12) * 3) : (u + upper_idx(j) * 3);
        p[j * 3] = src[0]; p[j * 3 + 1] = src[1]; p[j * 3 + 2] = src[2];
    }
}

// Error terms of ONE frame (Demo_test.py:64-69, 141-158) in the layout of MMEGO_SUMS_LEN: p = pred[63], g = target[63],
// u = upper_l[45], l = lower_l[24]; vals[46] is overwritten.
__device__ __forceinline__ void frame_metrics(const float* p, const float* g, const float* u, const float* l, float* vals) {
#pragma unroll
    for (int j = 0; j < 21; ++j) {
        const float dx = p[j * 3] - g[j * 3], dy = p[j * 3 + 1] - g[j * 3 + 1], dz = p[j * 3 + 2] - g[j * 3 + 2];
        vals[j] = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    float eu = 0.f, el = 0.f;
#pragma unroll
    for (int j = 0; j < 21; ++j) {
        if (j <= 12 || j == 16 || j == 20) {
            const int ui = upper_idx(j);
            const float dx = u[ui * 3] - g[j * 3], dy = u[ui * 3 + 1] - g[j * 3 + 1], dz = u[ui * 3 + 2] - g[j * 3 + 2];
            eu += sqrtf(dx * dx + dy * dy + dz * dz);
        }
        if (j >= 12 && j <= 19) {
            const int li = j - 12;
            const float dx = l[li * 3] - g[j * 3], dy = l[li * 3 + 1] - g[j * 3 + 1], dz = l[li * 3 + 2] - g[j * 3 + 2];
            el += sqrtf(dx * dx + dy * dy + dz * dz);
        }
    }
    vals[21] = eu;
    vals[22] = el;
#pragma unroll
    for (int i = 0; i < 20; ++i) {
        const int a = kSkelParent[i], b = kSkelChild[i];
        const float px = p[b * 3] - p[a * 3], py = p[b * 3 + 1] - p[a * 3 + 1], pz = p[b * 3 + 2] - p[a * 3 + 2];
        const float gx = g[b * 3] - g[a * 3], gy = g[b * 3 + 1] - g[a * 3 + 1], gz = g[b * 3 + 2] - g[a * 3 + 2];
        // torch cosine_similarity: dot / max(|p| * |g|, eps) with eps = 1e-8
        const float dot = px * gx + py * gy + pz * gz;
        const float den = fmaxf(sqrtf((px * px + py * py + pz * pz) * (gx * gx + gy * gy + gz * gz)), 1e-8f);
        float c = dot / den;
        c = fminf(fmaxf(c, -1.0f), 1.0f);
        vals[23 + i] = fabsf(acosf(c) / 3.14159265358f * 180.0f);
    }
    vals[43] = 1.0f;
    // L1 sums of the reference's eval_loss / eval_loss_l (Demo_test.py:141-147): lower joints and lower bone vectors
    float l1 = 0.f, l1b = 0.f;
#pragma unroll
    for (int k = 0; k < 24; ++k) l1 += fabsf(l[k] - g[36 + k]);
#pragma unroll
    for (int i = 14; i < 20; ++i) {
        const int a = kSkelParent[i] - 12, b = kSkelChild[i] - 12;
#pragma unroll
        for (int c = 0; c < 3; ++c) l1b += fabsf((l[b * 3 + c] - l[a * 3 + c]) - (g[36 + b * 3 + c] - g[36 + a * 3 + c]));
    }
    vals[44] = l1;
    vals[45] = l1b;
}

}  // namespace dec
}  // namespace mmego
