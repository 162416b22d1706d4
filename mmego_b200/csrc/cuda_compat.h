// cuda_compat.h -- one switch between the real CUDA toolchain (product build, nvcc, sm_100a) and the
// CPU execution-model emulator used by the "not gpu" test suite (tests/emul/, g++ -DMMEGO_EMUL).
#pragma once
namespace mmego { extern thread_local long long t_launches; }
#ifdef MMEGO_EMUL
#include "cuda_emul.h"
#define MMEGO_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
    do {                                                                              \
        ++mmego::t_launches;                                                          \
        emul::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); });        \
    } while (0)
#define MMEGO_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emul::dyn_smem)
#else
#include <cuda_runtime.h>
#define MMEGO_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
    do {                                                                              \
        ++mmego::t_launches;                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                   \
    } while (0)
#define MMEGO_DYN_SMEM(type, name)                                        \
    extern __shared__ __align__(1024) unsigned char mmego_dyn_smem_raw[]; \
    type* name = reinterpret_cast<type*>(mmego_dyn_smem_raw)
#endif
