// common.cuh -- shared device helpers for libcgd_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cgd_b200.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "libcgd_b200 must be compiled for sm_100a (-gencode arch=compute_100a,code=sm_100a)"
#endif
#endif

namespace cgd {

typedef __half act_t;  // activation / weight storage type of the networks (UMMA kind::f16, fp32 accumulate)

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
#define CGD_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      cgd::set_error(__VA_ARGS__);      \
      return -1;                        \
    }                                   \
  } while (0)
#define CGD_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      cgd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)
#define CGD_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      cgd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result broadcast to all threads; `red` = >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

// sigmoid with the SFU exp + approximate reciprocal (2 MUFU ops, ~2 ulp): these run per element in the HBM-bound
// normalisation kernels, where the IEEE divide sequence (~10 instructions) was a visible share of the issue slots
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
__device__ __forceinline__ float silu_f(float v) { return v * sigmoid_fast(v); }
// d silu(v)/dv
__device__ __forceinline__ float silu_grad_f(float v) {
  const float s = sigmoid_fast(v);
  return s * (1.f + v * (1.f - s));
}

struct __align__(16) half8 {
  __half2 a, b, c, d;
};
__device__ __forceinline__ void unpack8(const half8& h, float* f) {
  float2 t;
  t = __half22float2(h.a); f[0] = t.x; f[1] = t.y;
  t = __half22float2(h.b); f[2] = t.x; f[3] = t.y;
  t = __half22float2(h.c); f[4] = t.x; f[5] = t.y;
  t = __half22float2(h.d); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ half8 pack8(const float* f) {
  half8 h;
  h.a = __floats2half2_rn(f[0], f[1]);
  h.b = __floats2half2_rn(f[2], f[3]);
  h.c = __floats2half2_rn(f[4], f[5]);
  h.d = __floats2half2_rn(f[6], f[7]);
  return h;
}
__device__ __forceinline__ half8 ld8(const __half* p) { return *reinterpret_cast<const half8*>(p); }
__device__ __forceinline__ void st8(__half* p, const half8& v) { *reinterpret_cast<half8*>(p) = v; }

}  // namespace cgd
