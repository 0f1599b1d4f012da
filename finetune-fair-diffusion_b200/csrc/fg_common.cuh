// Shared device/host helpers for the fairguide sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/fairguide.h"

#define FG_NUM_SMS 148            // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define FG_LAUNCH_CHECK()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

#define FG_DISPATCH_DTYPE(dtype, T, ...)                                        \
    switch (dtype) {                                                            \
        case FG_F32:  { using T = float;         __VA_ARGS__; break; }          \
        case FG_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }          \
        case FG_F16:  { using T = __half;        __VA_ARGS__; break; }          \
        default: return FG_ERR_DTYPE;                                           \
    }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// round-trip through T: what a torch op that computes in fp32 and stores T leaves behind
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f32<T>(from_f32<T>(v)); }

static inline size_t fg_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline cudaStream_t fg_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
