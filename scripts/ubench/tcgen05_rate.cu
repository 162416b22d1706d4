// Microbenchmark: tcgen05.mma issue rates on sm_100a for kind::f16 (fp16 in, fp32 accumulate), kind::i8 (int8 in, int32
// accumulate) and kind::f8f6f4 (e4m3 in, fp32 accumulate), M=128 N=256 per SM (cta_group::1), operands in shared memory.
// Question behind it: could the two cross terms of the split-precision LSTM GEMM (a_hi*w_lo + a_lo*w_hi) run as int8
// MMAs at twice the fp16 rate?  Modes:
//   0  fp16 only                    (K = 16 per MMA)
//   1  int8 only                    (K = 32 per MMA)
//   2  e4m3 only                    (K = 32 per MMA)
//   3  today's scheme per 64-wide K block: 12 fp16 MMAs into one fp32 accumulator
//   4  proposed scheme per 64-wide K block: 4 fp16 MMAs (fp32 accumulator) + 4 int8 MMAs (int32 accumulator)
// Every mode also checks one accumulator element against the value expected from the constant operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_rate tcgen05_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../mmego_b200/csrc/tc_common.cuh"

using namespace mmego::tc;

__device__ __forceinline__ void mma_i8_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_f8_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

constexpr int BM = 128, BN = 256;
constexpr uint32_t IDESC_F16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);               // D fp32, A/B fp16
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // D s32, A/B int8
constexpr uint32_t IDESC_F8 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);                // D fp32, A/B e4m3

// smem: A16 [128 rows x 128 B], B16 [256 x 128 B], A8 [128 x 128 B], B8 [256 x 128 B]; every tile 1024-byte aligned
constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
constexpr int SMEM = 2 * (A_BYTES + B_BYTES) + 1024 + 64;

__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int iters, float* out_f, int* out_i, int random_data) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* a16 = smem;
    uint8_t* b16 = a16 + A_BYTES;
    uint8_t* a8 = b16 + B_BYTES;
    uint8_t* b8 = a8 + A_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b8 + B_BYTES);
    uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    // operand fill: constants for the check (fp16 1.0 x 0.5, int8 3 x -2, e4m3 1.0 x 0.5), or pseudo-random "realistic" bits
    uint32_t rng = 0x9E3779B9u * (blockIdx.x * 128 + tid + 1);
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; };
    for (int i = tid; i < (A_BYTES + B_BYTES) / 2; i += 128) {
        const bool isA = i < A_BYTES / 2;
        __half v = __float2half(isA ? 1.0f : 0.5f);
        if (random_data) v = __float2half(((int)(next() & 0xFFFF) - 32768) / 32768.0f);
        // operand-entropy experiment: random_data 2 / 3 = B operand with its low 5 / 8 mantissa bits cleared, 4 = both
        // operands with 5 cleared (does the tensor core draw less power on "short" residual planes?)
        unsigned short bits = *reinterpret_cast<unsigned short*>(&v);
        if ((random_data == 2 && !isA) || random_data == 4) bits &= 0xFFE0;
        if (random_data == 3 && !isA) bits &= 0xFF00;
        v = *reinterpret_cast<__half*>(&bits);
        reinterpret_cast<__half*>(a16)[i] = v;
    }
    for (int i = tid; i < A_BYTES + B_BYTES; i += 128) {
        const bool isA = i < A_BYTES;
        uint8_t v;
        if (mode == 2) {
            v = isA ? 0x38 : 0x30;                                   // e4m3 1.0, 0.5
            if (random_data) { v = next() & 0xFF; if ((v & 0x7F) == 0x7F) v ^= 1; }
        } else {
            v = isA ? (uint8_t)3 : (uint8_t)(-2);
            if (random_data) v = next() & 0xFF;
        }
        a8[i] = v;
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(holder, 512);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;

    if (tid == 0) {
        const uint32_t A16 = smem_u32(a16), B16 = smem_u32(b16), A8 = smem_u32(a8), B8 = smem_u32(b8);
        for (int it = 0; it < iters; ++it) {
            const uint32_t first = it > 0;
            if (mode == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_f16_ss(tmem, make_sw128_kmajor_desc(A16 + k * 32), make_sw128_kmajor_desc(B16 + k * 32), IDESC_F16, first | (k > 0));
            } else if (mode == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_i8_ss(tmem + 256, make_sw128_kmajor_desc(A8 + k * 32), make_sw128_kmajor_desc(B8 + k * 32), IDESC_I8, first | (k > 0));
            } else if (mode == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_f8_ss(tmem, make_sw128_kmajor_desc(A8 + k * 32), make_sw128_kmajor_desc(B8 + k * 32), IDESC_F8, first | (k > 0));
            } else if (mode == 3) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass)
                        mma_f16_ss(tmem, make_sw128_kmajor_desc(A16 + k * 32), make_sw128_kmajor_desc(B16 + k * 32), IDESC_F16, first | (k > 0) | (pass > 0));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mma_f16_ss(tmem, make_sw128_kmajor_desc(A16 + k * 32), make_sw128_kmajor_desc(B16 + k * 32), IDESC_F16, first | (k > 0));
                    // 64 K elements of int8 = bytes 0..63 of the 128-byte rows: two K=32 MMAs per cross term
                    if (k < 2) {
                        mma_i8_ss(tmem + 256, make_sw128_kmajor_desc(A8 + k * 32), make_sw128_kmajor_desc(B8 + k * 32), IDESC_I8, first | (k > 0));
                        mma_i8_ss(tmem + 256, make_sw128_kmajor_desc(A8 + 64 + k * 32), make_sw128_kmajor_desc(B8 + 64 + k * 32), IDESC_I8, 1);
                    }
                }
            }
        }
        mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    // element (row = tid, column 0) of both accumulators
    uint32_t r[8];
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    tmem_ld_x8(lane_addr, r);
    tmem_ld_wait();
    out_f[blockIdx.x * 128 + tid] = __uint_as_float(r[0]);
    tmem_ld_x8(lane_addr + 256, r);
    tmem_ld_wait();
    out_i[blockIdx.x * 128 + tid] = (int)r[0];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 1.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    float* of;
    int* oi;
    cudaMalloc(&of, sms * 128 * 4);
    cudaMalloc(&oi, sms * 128 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[5] = {"fp16 only", "int8 only", "e4m3 only", "today: 12 fp16 MMAs per K block", "proposed: 4 fp16 + 4 int8 MMAs per K block"};
    // correctness of the descriptors with constant operands (4 MMAs per iteration, one iteration)
    for (int mode = 0; mode < 5; ++mode) {
        rate_kernel<<<sms, 128, SMEM>>>(mode, 1, of, oi, 0);
        cudaError_t err = cudaDeviceSynchronize();
        float hf[128];
        int hi[128];
        cudaMemcpy(hf, of, sizeof(hf), cudaMemcpyDeviceToHost);
        cudaMemcpy(hi, oi, sizeof(hi), cudaMemcpyDeviceToHost);
        printf("check mode %d (%s): %s  fp32 acc[0][0]=%g acc[127][0]=%g  int acc[0][0]=%d acc[127][0]=%d\n", mode, names[mode],
               cudaGetErrorString(err), hf[0], hf[127], hi[0], hi[127]);
    }
    printf("expected: fp16 64*0.5 = 32 (mode 3: 96); int8 K=128 per iteration in modes 1 (3*-2*128 = -768), mode 4 same -768; e4m3 128*0.5 = 64\n");
    const int max_rd = argc > 2 ? atoi(argv[2]) : 1;
    for (int random_data = 0; random_data <= max_rd; ++random_data)
        for (int mode = 0; mode < 5; ++mode) {
            if (random_data >= 2 && mode != 0) continue;
            int iters = 20000;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {          // first pass calibrates the iteration count to `seconds`
                cudaEventRecord(e0);
                rate_kernel<<<sms, 128, SMEM>>>(mode, iters, of, oi, random_data);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                if (err != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(err)); return 1; }
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep == 0) iters = (int)(iters * seconds * 1e3 / ms);
            }
            const double per_iter_ns = ms * 1e6 / iters;
            // MMA instruction counts per iteration and their MAC volume
            double macs = 0;
            if (mode == 0) macs = 4.0 * BM * BN * 16;
            if (mode == 1 || mode == 2) macs = 4.0 * BM * BN * 32;
            if (mode == 3) macs = 12.0 * BM * BN * 16;
            if (mode == 4) macs = 4.0 * BM * BN * 16 + 4.0 * BM * BN * 32;
            printf("%s data, mode %d (%s): %.1f ns per iteration per SM, %.0f T(FL)OP/s over %d SMs (%.2f s)\n",
                   random_data == 0 ? "constant" : random_data == 1 ? "random" : random_data == 2 ? "random, B low 5 mantissa bits 0" : random_data == 3 ? "random, B low 8 mantissa bits 0" : "random, A and B low 5 bits 0", mode, names[mode], per_iter_ns, 2 * macs * sms / (per_iter_ns * 1e-9) / 1e12, sms,
                   ms / 1e3);
        }
    return 0;
}
