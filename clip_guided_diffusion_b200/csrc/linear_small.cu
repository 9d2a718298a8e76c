// linear_small.cu -- small-M fp32 linear layers: y[M,N] (=|+=) act(x[M,K]) @ W[N,K]^T + b.
//
// Replaces the [3P] UNet time_embed MLP and the ~30 per-ResBlock emb_layers Linears (x-independent, forward
// only) and the [3P] CLIP head ln_post(cls) @ proj with its transpose for the backward (K16).
// M is the batch (1..64 rows) so these are weight-streaming GEMVs: one warp per output column, the weight
// row read once with 128-bit loads, M accumulators per lane, x staged in shared memory.
#include <algorithm>

#include "common.cuh"
#include "pdl.cuh"
#include "ops.cuh"

namespace cgd {

constexpr int LS_MT = 8;  // rows of x per block

template <typename XT>
__global__ void linear_small_kernel(const XT* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ bias, void* __restrict__ y,
                                    int M, int K, int N, int64_t ldx, int64_t ldy, int silu_in, int accumulate, int y_half) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float xs[];  // [LS_MT][K]
  const int m0 = blockIdx.y * LS_MT, mt = min(LS_MT, M - m0);
  for (int i = threadIdx.x; i < mt * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    float v = (float)x[(int64_t)(m0 + r) * ldx + c];
    xs[r * K + c] = silu_in ? silu_f(v) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int n = blockIdx.x * nw + wid; n < N; n += gridDim.x * nw) {
    const float* wr = Wt + (int64_t)n * K;
    float acc[LS_MT];
#pragma unroll
    for (int r = 0; r < LS_MT; ++r) acc[r] = 0.f;
    for (int c = lane * 4; c < K; c += 128) {
      const float4 w4 = *reinterpret_cast<const float4*>(wr + c);
#pragma unroll
      for (int r = 0; r < LS_MT; ++r) {
        if (r < mt) {
          const float4 x4 = *reinterpret_cast<const float4*>(xs + r * K + c);
          acc[r] = fmaf(w4.x, x4.x, fmaf(w4.y, x4.y, fmaf(w4.z, x4.z, fmaf(w4.w, x4.w, acc[r]))));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < LS_MT; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float bv = bias ? bias[n] : 0.f;
      for (int r = 0; r < mt; ++r) {
        const int64_t o = (int64_t)(m0 + r) * ldy + n;
        float v = acc[r] + bv;
        if (y_half) {
          __half* yh = reinterpret_cast<__half*>(y);
          if (accumulate) v += __half2float(yh[o]);
          yh[o] = __float2half_rn(v);
        } else {
          float* yf = reinterpret_cast<float*>(y);
          if (accumulate) v += yf[o];
          yf[o] = v;
        }
      }
    }
  }
}

int launch_linear_small(const CgdOp& op, cudaStream_t st) {
  const int64_t M = op.i[0], K = op.i[1], N = op.i[2], ldx = op.i[3], ldy = op.i[4];
  CGD_CHECK_ARG(M > 0 && K > 0 && K % 4 == 0 && N > 0 && op.p[0] && op.p[1] && op.p[3], "linear_small: bad args (M=%lld K=%lld N=%lld)",
                (long long)M, (long long)K, (long long)N);
  CGD_CHECK_ARG((size_t)LS_MT * K * sizeof(float) <= 96 * 1024, "linear_small: K=%lld too large", (long long)K);
  const int smem = (int)(LS_MT * K * sizeof(float));
  const int silu = op.flags & 1, acc = (op.flags & 2) ? 1 : 0, xh = (op.flags & 4) ? 1 : 0, yh = (op.flags & 8) ? 1 : 0;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(N, 8), 148 * 4), (unsigned)ceil_div(M, LS_MT));
  if (xh) {
    static bool set = false;
    if (!set) { CGD_CUDA(cudaFuncSetAttribute(linear_small_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); set = true; }
    CGD_CUDA(launch_pdl(linear_small_kernel<__half>, dim3(grid), dim3(256), smem, st, (const __half*)op.p[0], (const float*)op.p[1], (const float*)op.p[2], op.p[3], (int)M, (int)K,
                                                        (int)N, ldx, ldy, silu, acc, yh));
  } else {
    static bool set = false;
    if (!set) { CGD_CUDA(cudaFuncSetAttribute(linear_small_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); set = true; }
    CGD_CUDA(launch_pdl(linear_small_kernel<float>, dim3(grid), dim3(256), smem, st, (const float*)op.p[0], (const float*)op.p[1], (const float*)op.p[2], op.p[3], (int)M, (int)K,
                                                       (int)N, ldx, ldy, silu, acc, yh));
  }
  CGD_LAUNCH_CHECK();
  return 0;
}

}  // namespace cgd
