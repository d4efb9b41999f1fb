// common.cuh — shared helpers for libvtoonify_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/vtoonify_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvtoonify_b200 is written for sm_100a (B200) only"
#endif

int vt_set_error(const char* fmt, ...);
void vt_count_launch(int n);

#define VT_CHECK(cond, ...)                          \
  do {                                               \
    if (!(cond)) return vt_set_error(__VA_ARGS__);   \
  } while (0)

#define VT_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return vt_set_error("%s:%d %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

#define VT_LAUNCH_CHECK()                                                                 \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return vt_set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    vt_count_launch(1);                                                                   \
  } while (0)

static inline __host__ __device__ int64_t vt_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Round fp32 to TF32 (round-to-nearest, ties away) keeping an fp32 container. The tensor core
// reads only the top 19 bits of the container, so pre-rounding in the producer makes the
// truncation unbiased.
__device__ __forceinline__ float vt_round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float vt_lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

int vt_num_sms();
