// common.cuh — shared helpers for libgraingnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/graingnn_b200.h"

#define GG_STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define GG_LAUNCH_OK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

static inline bool gg_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 128-bit read-only load (LDG.E.128.CONSTANT); rows must be 16-byte aligned.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
