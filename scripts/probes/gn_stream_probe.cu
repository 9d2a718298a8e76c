// Stand-alone probe (NOT part of the library): what bounds a GroupNorm-apply-like stream y = silu(x * A[c] + B[c]) over a
// [65536, 256] fp16 tensor on B200?  Variants isolate the load path, the conversions, the SiLU and the packed-f32x2 arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/gn_probe scripts/probes/gn_stream_probe.cu && /tmp/gn_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint4 ld_nc(const __half* p) { uint4 v; asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd;}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// MODE: 0 copy, 1 affine scalar, 2 affine + silu(tanh) scalar, 3 affine + silu(ex2, rcp) scalar, 4 packed f32x2 + silu(tanh) (library form),
//       5 = 2 with plain ld.global (no .nc), 6 = 2 with two rows per trip
template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ A, const float* __restrict__ B, int HW, int C) {
  const int V = C / 8, RP = 256 / V, col = threadIdx.x % V, slot = threadIdx.x / V;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = A[col * 8 + j]; b[j] = B[col * 8 + j]; }
  const int st = gridDim.x * RP;
  auto body = [&](uint4 v, int r) {
    if (MODE == 0) { *reinterpret_cast<uint4*>(y + (int64_t)r * C + col * 8) = v; return; }
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      float2 t;
      if (MODE == 4) {
        t = fma2(f, make_float2(a[2 * j], a[2 * j + 1]), make_float2(b[2 * j], b[2 * j + 1]));
        const float2 h = mul2(t, make_float2(0.5f, 0.5f));
        t = fma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
      } else {
        t.x = fmaf(f.x, a[2 * j], b[2 * j]);
        t.y = fmaf(f.y, a[2 * j + 1], b[2 * j + 1]);
        if (MODE == 2 || MODE == 5 || MODE == 6) {
          const float hx = 0.5f * t.x, hy = 0.5f * t.y;
          t.x = fmaf(hx, tanh_approx(hx), hx);
          t.y = fmaf(hy, tanh_approx(hy), hy);
        } else if (MODE == 3) {
          t.x = t.x * rcpf(1.f + ex2f(-1.4426950408889634f * t.x));
          t.y = t.y * rcpf(1.f + ex2f(-1.4426950408889634f * t.y));
        }
      }
      const __half2 h2 = __floats2half2_rn(t.x, t.y);
      o[j] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(y + (int64_t)r * C + col * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  if (slot >= RP) return;
  int r = blockIdx.x * RP + slot;
  if (MODE == 6) {
    for (; r + st < HW; r += 2 * st) {
      const uint4 v0 = ld_nc(x + (int64_t)r * C + col * 8), v1 = ld_nc(x + (int64_t)(r + st) * C + col * 8);
      body(v0, r);
      body(v1, r + st);
    }
  }
  for (; r < HW; r += st) {
    uint4 v;
    if (MODE == 5) v = *reinterpret_cast<const uint4*>(x + (int64_t)r * C + col * 8);
    else v = ld_nc(x + (int64_t)r * C + col * 8);
    body(v, r);
  }
}

template <int MODE, int MINB>
static float run(const char* name, __half** xs, __half** ys, int nbuf, const float* A, const float* B, int HW, int C, int grid) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < nbuf; ++i) k<MODE, MINB><<<grid, 256>>>(xs[i], ys[i], A, B, HW, C);
  cudaDeviceSynchronize();
  const int reps = 5;
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r)
    for (int i = 0; i < nbuf; ++i) k<MODE, MINB><<<grid, 256>>>(xs[i], ys[i], A, B, HW, C);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const float us = ms * 1e3f / (reps * nbuf);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<MODE, MINB>, 256, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k<MODE, MINB>);
  printf("%-44s grid %5d  %6.1f us  %6.0f GB/s   regs %3d  CTAs/SM %d\n", name, grid, us, 2.0 * HW * C * 2 / us / 1e3, fa.numRegs, nb);
  return us;
}

int main() {
  const int HW = 65536, C = 256, nbuf = 6;  // 6 x (33.5 in + 33.5 out) MB = 402 MB > 126 MB L2
  __half *xs[nbuf], *ys[nbuf];
  for (int i = 0; i < nbuf; ++i) {
    CK(cudaMalloc(&xs[i], (size_t)HW * C * 2));
    CK(cudaMalloc(&ys[i], (size_t)HW * C * 2));
    CK(cudaMemset(xs[i], 0x3c, (size_t)HW * C * 2));
  }
  float *A, *B;
  CK(cudaMalloc(&A, C * 4));
  CK(cudaMalloc(&B, C * 4));
  CK(cudaMemset(A, 0, C * 4));
  CK(cudaMemset(B, 0, C * 4));
  for (int grid : {1184, 2368, 592}) {
    printf("--- grid %d\n", grid);
    run<0, 1>("copy", xs, ys, nbuf, A, B, HW, C, grid);
    run<1, 1>("affine (scalar fma)", xs, ys, nbuf, A, B, HW, C, grid);
    run<2, 1>("affine + silu tanh (scalar)", xs, ys, nbuf, A, B, HW, C, grid);
    run<3, 1>("affine + silu ex2+rcp (scalar)", xs, ys, nbuf, A, B, HW, C, grid);
    run<4, 1>("affine + silu tanh, packed f32x2 (library)", xs, ys, nbuf, A, B, HW, C, grid);
    run<5, 1>("affine + silu tanh (scalar), plain ld", xs, ys, nbuf, A, B, HW, C, grid);
    run<6, 1>("affine + silu tanh (scalar), 2 rows / trip", xs, ys, nbuf, A, B, HW, C, grid);
    run<2, 8>("affine + silu tanh (scalar), minblocks 8", xs, ys, nbuf, A, B, HW, C, grid);
  }
  return 0;
}
