// Shared host/device helpers for libnbk_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/nbk_b200.h"

#define NBK_SM_COUNT 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void nbk_set_error(const char *fmt, ...);
void nbk_count_launch(int n = 1);

#define NBK_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            nbk_set_error(__VA_ARGS__);          \
            return NBK_ERR_ARG;                  \
        }                                        \
    } while (0)

#define NBK_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            nbk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return NBK_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define NBK_LAUNCHED()                                                                    \
    do {                                                                                  \
        nbk_count_launch();                                                               \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            nbk_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return NBK_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

static inline int nbk_grid_for(int64_t n, int block, int ctas_per_sm) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)NBK_SM_COUNT * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// integer frequency label: j -> j if j < ceil(N/2) else j - N  (Nyquist negative, meshtools.py:150-153)
__host__ __device__ __forceinline__ int nbk_freq(int j, int N) { return (j >= (N + 1) / 2) ? j - N : j; }
