// cuda_emul.h -- TEST INFRASTRUCTURE ONLY.  A minimal CUDA execution-model emulator so that the
// framework's plain-CUDA kernels (everything that is not inline PTX) can be compiled with g++ and
// run on the build container's CPU against the oracle before GPU minutes are spent.
//
// * one GPU thread = one ucontext fiber; __syncthreads / warp shuffles are cooperative barriers
// * blocks are distributed over a few OS worker threads; __shared__ maps to static thread_local
// * the runtime API subset used by the library (malloc/memcpy/memset/streams/events) is synchronous
//
// The product library (libmmego_b200.so, built by nvcc) never includes this file; the loader in
// mmego_b200/_capi.py never loads the emulated build.  It is compiled only by tests/emul/build_emul.py.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

namespace emul {
struct Ctx {
    dim3 tid_, bid_, bdim_, gdim_;
};
extern thread_local Ctx ctx;
extern thread_local unsigned char* dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void syncthreads();
void warp_exchange(const void* in, void* out, size_t bytes, int src_lane);
void warp_barrier();
void warp_allgather(const void* in, void* out_all, size_t bytes);   // out_all: [32][bytes], bytes <= 32
int lane_id();
}  // namespace emul

#define threadIdx (emul::ctx.tid_)
#define blockIdx (emul::ctx.bid_)
#define blockDim (emul::ctx.bdim_)
#define gridDim (emul::ctx.gdim_)
#define warpSize 32

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

static inline void __syncthreads() { emul::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = emul::lane_id();
    int base = lane & ~(width - 1);
    T r;
    emul::warp_exchange(&v, &r, sizeof(T), base + (src & (width - 1)));
    return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
    (void)width;
    T r;
    emul::warp_exchange(&v, &r, sizeof(T), emul::lane_id() ^ mask);
    return r;
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    int lane = emul::lane_id();
    int src = lane + (int)delta;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    T r;
    emul::warp_exchange(&v, &r, sizeof(T), src);
    return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    int lane = emul::lane_id();
    int src = lane - (int)delta;
    if (src < (lane & ~(width - 1))) src = lane;
    T r;
    emul::warp_exchange(&v, &r, sizeof(T), src);
    return r;
}

template <class T>
static inline T __ldg(const T* p) { return *p; }

static inline float atomicAdd(float* p, float v) {
    auto* a = reinterpret_cast<std::atomic<uint32_t>*>(p);
    uint32_t old = a->load(), nw;
    float f;
    do {
        std::memcpy(&f, &old, 4);
        float r = f + v;
        std::memcpy(&nw, &r, 4);
    } while (!a->compare_exchange_weak(old, nw));
    return f;
}
static inline double atomicAdd(double* p, double v) {
    auto* a = reinterpret_cast<std::atomic<uint64_t>*>(p);
    uint64_t old = a->load(), nw;
    double f;
    do {
        std::memcpy(&f, &old, 8);
        double r = f + v;
        std::memcpy(&nw, &r, 8);
    } while (!a->compare_exchange_weak(old, nw));
    return f;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
#define __expf(x) std::exp((float)(x))
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
using std::max;
using std::min;

// ---- runtime API subset (synchronous) --------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct cudaDeviceProp { int major, minor, multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
template <class T>
static inline cudaError_t cudaMemcpyToSymbolAsync(T& sym, const void* s, size_t n, size_t off, cudaMemcpyKind, cudaStream_t = nullptr) {
    std::memcpy(reinterpret_cast<char*>(&sym) + off, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    p->major = 10; p->minor = 0; p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024;
    std::snprintf(p->name, sizeof(p->name), "emulated-sm100");
    return cudaSuccess;
}
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
