// Shared helpers for libmvsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

#include "mvsb200.h"

namespace mvsb200 {

void set_error(const char *fmt, ...);
void clear_error();

inline int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MVSB200_E_CUDA;
    }
    return MVSB200_OK;
}

// Opt a kernel into `bytes` of dynamic shared memory.  The attribute is per (kernel, device): remembered in a small
// table so that a process driving several GPUs (nn.DataParallel worker threads) configures each of them.
// (Keyed by the kernel's ADDRESS: all instantiations of a kernel template share one function-pointer type.)
int ensure_dynamic_smem_impl(const void *kernel, size_t bytes, const char *what);
template <typename K> inline int ensure_dynamic_smem(K kernel, size_t bytes, const char *what)
{
    return ensure_dynamic_smem_impl(reinterpret_cast<const void *>(kernel), bytes, what);
}

#define MVSB200_REQUIRE(cond, ...)          \
    do {                                    \
        if (!(cond)) {                      \
            mvsb200::set_error(__VA_ARGS__); \
            return MVSB200_E_INVALID;       \
        }                                   \
    } while (0)

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// streaming store: the written volume is consumed by a later kernel, not by this one
__device__ __forceinline__ void st4_stream(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// Per-voxel depth hypothesis for the four depth modes of mvsb200.h.
__device__ __forceinline__ float hypothesis(int mode, const float *depth, float interval, int b, int d, int D,
                                            long long hw, long long pix)
{
    switch (mode) {
    case MVSB200_DEPTH_VALUES: return __ldg(depth + (long long)b * D + d);
    case MVSB200_DEPTH_VOLUME: return __ldg(depth + ((long long)b * D + d) * hw + pix);
    case MVSB200_DEPTH_START: return __ldg(depth + b) + interval * (float)d;
    default: return __ldg(depth + (long long)b * hw + pix) + interval * (float)d;
    }
}

}  // namespace mvsb200
