// norm_grid.cu -- GroupNorm(32) (+SiLU, +scale/shift) forward and input-gradient for LARGE activations as ONE persistent
// launch: statistics pass, grid-wide barrier, apply pass.
//
// Why (measured on B200, profiles/r01_gn_microbench_v1.txt): at the 128x128 / 256x256 levels the two-launch kernels of norm.cu
// reach 0.9 - 1.7 TB/s of the 6.5 TB/s copy bandwidth -- two blocks of 256 threads per SM keep ~5 MB in flight, the grids end
// in half-empty waves, and the second launch waits for a last-block fold.  Here one CTA per SM (512 threads, 8 independent
// 128-bit loads per thread in flight = 9.7 MB over the chip) owns a contiguous pixel range of one image across ALL channels
// (full 128-byte lines), writes its per-group partial sums, meets the other CTAs at a generation-counted grid barrier (every
// CTA is resident: grid <= SM count, one CTA per SM), folds the partials itself (fixed order, fp64: bit-identical in every
// CTA) and applies the normalisation to the same pixel range, whose second read is an L2 hit (the tensors are 8 - 67 MB,
// the L2 is 126 MB).  HBM traffic: x once + y once (forward), dy + x once + dx (backward).
//
// Replaces [3P] guided-diffusion GroupNorm32 + SiLU + scale-shift and their autograd.
#include "common.cuh"
#include "ops.cuh"
#include "pdl.cuh"

namespace cgd {

constexpr int kGngThreads = 512;
constexpr int kGngUnroll = 8;

// Generation-counted barrier over all CTAs of the grid.  bar[0] = arrival count (returns to 0), bar[1] = generation (only
// ever incremented), so the buffer needs no reset between launches or CUDA-graph replays.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* vgen = bar + 1;
    const unsigned int gen = *vgen;  // read before arriving: the generation cannot advance until this CTA has arrived
    __threadfence();                 // cumulative: publishes the partials written by this CTA's other threads (bar.sync above)
    if (atomicAdd(bar, 1u) == nblocks - 1u) {
      atomicExch(bar, 0u);
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*vgen == gen) __nanosleep(40);
    }
    __threadfence();
  }
  __syncthreads();
}

// 8 channel sums per thread -> 32 group sums of this CTA.  Shared layout [pixel lane][channel]; thread (g = t/16, part = t%16)
// adds elements part, part+16, ... of group g's PP * cpg values, then the 16 parts are folded by shuffles (fixed order).
__device__ __forceinline__ void gng_block_reduce(const float* s, const float* q, int C, int col, int pl, int PP, bool active,
                                                 float* red_s, float* red_q, float* gs, float* gq) {
  __syncthreads();
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red_s[pl * C + col * 8 + j] = s[j];
      red_q[pl * C + col * 8 + j] = q[j];
    }
  }
  __syncthreads();
  const int cpg = C / 32, g = threadIdx.x >> 4, part = threadIdx.x & 15;
  float as = 0.f, aq = 0.f;
  const int n = PP * cpg;
  for (int e = part; e < n; e += 16) {
    const int l = e / cpg, c = e - l * cpg;
    as += red_s[l * C + g * cpg + c];
    aq += red_q[l * C + g * cpg + c];
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    as += __shfl_xor_sync(0xffffffffu, as, o);
    aq += __shfl_xor_sync(0xffffffffu, aq, o);
  }
  if (part == 0) {
    gs[g] = as;
    gq[g] = aq;
  }
  __syncthreads();
}

// Fold the Gn per-CTA partials of image n (layout [Gn][32][2]) in a fixed order, fp64: thread (g, part) takes CTAs part,
// part+16, ...; the 16 parts are folded by shuffles.  Results for all 32 groups land in shared memory.
__device__ __forceinline__ void gng_fold(const float* part_n, int Gn, double* out_a, double* out_b) {
  const int g = threadIdx.x >> 4, part = threadIdx.x & 15;
  double da = 0.0, db = 0.0;
  for (int j = part; j < Gn; j += 16) {
    const float2 v = __ldcg(reinterpret_cast<const float2*>(part_n + ((int64_t)j * 32 + g) * 2));
    da += (double)v.x;
    db += (double)v.y;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    da += __shfl_xor_sync(0xffffffffu, da, o);
    db += __shfl_xor_sync(0xffffffffu, db, o);
  }
  if (part == 0) {
    out_a[g] = da;
    out_b[g] = db;
  }
  __syncthreads();
}

struct GngGeom {
  int V, PP, threads;
};
__host__ __device__ inline GngGeom gng_geom(int C) {
  GngGeom g;
  g.V = C / 8;
  g.PP = kGngThreads / g.V;
  g.threads = kGngThreads;
  return g;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kGngThreads, 1)
gn_fwd_grid_kernel(const __half* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ emb, __half* __restrict__ y, float* __restrict__ stats, float* __restrict__ partials,
                   unsigned int* __restrict__ bar, int HW, int C, int64_t ldx, int64_t ldy, int Gn, float eps, int silu) {
  __shared__ float red_s[kGngThreads * 8], red_q[kGngThreads * 8];
  __shared__ float gs[32], gq[32];
  __shared__ double fa[32], fb[32];
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.x / Gn, chunk = blockIdx.x % Gn;
  const int V = C / 8, PP = kGngThreads / V;
  const int col = threadIdx.x % V, pl = threadIdx.x / V;
  const bool active = pl < PP;
  const int cpg = C / 32;
  const int ppc = (HW + Gn - 1) / Gn;
  const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
  float ga[8], be[8];
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ga[j] = gamma[col * 8 + j];
      be[j] = beta[col * 8 + j];
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  const __half* xb = x + (int64_t)n * HW * ldx + col * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active) {
    for (int p = p0 + pl; p < p1; p += kGngUnroll * PP) {
      half8 raw[kGngUnroll];
#pragma unroll
      for (int u = 0; u < kGngUnroll; ++u)
        if (p + u * PP < p1) raw[u] = ld8(xb + (int64_t)(p + u * PP) * ldx);
#pragma unroll
      for (int u = 0; u < kGngUnroll; ++u) {
        if (p + u * PP < p1) {
          float v[8];
          unpack8(raw[u], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += v[j];
            q[j] = fmaf(v[j], v[j], q[j]);
          }
        }
      }
    }
  }
  gng_block_reduce(s, q, C, col, pl, PP, active, red_s, red_q, gs, gq);
  if (threadIdx.x < 32) {
    float* o = partials + (((int64_t)n * Gn + chunk) * 32 + threadIdx.x) * 2;
    o[0] = gs[threadIdx.x];
    o[1] = gq[threadIdx.x];
  }
  grid_barrier(bar, gridDim.x);
  gng_fold(partials + (int64_t)n * Gn * 64, Gn, fa, fb);
  if (threadIdx.x < 32) {
    const double m = (double)cpg * (double)HW;
    const double mean = fa[threadIdx.x] / m;
    double var = fb[threadIdx.x] / m - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rs = (float)(1.0 / sqrt(var + (double)eps));
    s_mean[threadIdx.x] = mu;
    s_rstd[threadIdx.x] = rs;
    if (chunk == 0) {
      stats[((int64_t)n * 32 + threadIdx.x) * 2 + 0] = mu;
      stats[((int64_t)n * 32 + threadIdx.x) * 2 + 1] = rs;
    }
  }
  __syncthreads();
  if (!active) return;
  float A[8], Bc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = col * 8 + j, g = c / cpg;
    const float mu = s_mean[g], rs = s_rstd[g];
    float sc1 = 1.f, sh = 0.f;
    if (emb) {
      sc1 = 1.f + emb[(int64_t)n * 2 * C + c];
      sh = emb[(int64_t)n * 2 * C + C + c];
    }
    A[j] = rs * ga[j] * sc1;
    Bc[j] = (be[j] - mu * rs * ga[j]) * sc1 + sh;
  }
  __half* yb = y + (int64_t)n * HW * ldy + col * 8;
  for (int p = p0 + pl; p < p1; p += kGngUnroll * PP) {
    half8 raw[kGngUnroll];
#pragma unroll
    for (int u = 0; u < kGngUnroll; ++u)
      if (p + u * PP < p1) raw[u] = ld8(xb + (int64_t)(p + u * PP) * ldx);
#pragma unroll
    for (int u = 0; u < kGngUnroll; ++u) {
      if (p + u * PP < p1) {
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(v[j], A[j], Bc[j]);
          v[j] = silu ? silu_f(t) : t;
        }
        st8(yb + (int64_t)(p + u * PP) * ldy, pack8(v));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kGngUnrollB = 4;

__global__ void __launch_bounds__(kGngThreads, 1)
gn_bwd_grid_kernel(const __half* __restrict__ dy, const __half* __restrict__ x, const float* __restrict__ stats,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ emb,
                   __half* __restrict__ dx, float* __restrict__ partials, unsigned int* __restrict__ bar, int HW, int C, int64_t ld_dy,
                   int64_t ldx, int64_t ld_dx, int Gn, int silu, int accumulate) {
  __shared__ float red_s[kGngThreads * 8], red_q[kGngThreads * 8];
  __shared__ float gs[32], gq[32];
  __shared__ double fa[32], fb[32];
  __shared__ float s_m1[32], s_m2[32];
  const int n = blockIdx.x / Gn, chunk = blockIdx.x % Gn;
  const int V = C / 8, PP = kGngThreads / V;
  const int col = threadIdx.x % V, pl = threadIdx.x / V;
  const bool active = pl < PP;
  const int cpg = C / 32;
  const int ppc = (HW + Gn - 1) / Gn;
  const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
  float G[8], Bc[8];
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      G[j] = gamma[col * 8 + j];
      Bc[j] = beta[col * 8 + j];
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  float mu[8], rs[8];
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col * 8 + j, g = c / cpg;
      mu[j] = stats[((int64_t)n * 32 + g) * 2];
      rs[j] = stats[((int64_t)n * 32 + g) * 2 + 1];
      float sc1 = 1.f, sh = 0.f;
      if (emb) {
        sc1 = 1.f + emb[(int64_t)n * 2 * C + c];
        sh = emb[(int64_t)n * 2 * C + C + c];
      }
      const float ga = G[j];
      G[j] = ga * sc1;                                  // d v / d xhat
      Bc[j] = (Bc[j] - mu[j] * rs[j] * ga) * sc1 + sh;  // v = x * (rs * G) + Bc
    }
  }
  const __half* xb = x + (int64_t)n * HW * ldx + col * 8;
  const __half* db = dy + (int64_t)n * HW * ld_dy + col * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active) {
    for (int p = p0 + pl; p < p1; p += kGngUnrollB * PP) {
      half8 rx[kGngUnrollB], rd[kGngUnrollB];
#pragma unroll
      for (int u = 0; u < kGngUnrollB; ++u)
        if (p + u * PP < p1) {
          rx[u] = ld8(xb + (int64_t)(p + u * PP) * ldx);
          rd[u] = ld8(db + (int64_t)(p + u * PP) * ld_dy);
        }
#pragma unroll
      for (int u = 0; u < kGngUnrollB; ++u) {
        if (p + u * PP < p1) {
          float v[8], d[8];
          unpack8(rx[u], v);
          unpack8(rd[u], d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float dv = d[j];
            if (silu) dv *= silu_grad_f(fmaf(v[j], rs[j] * G[j], Bc[j]));
            const float dxh = dv * G[j];
            const float xh = (v[j] - mu[j]) * rs[j];
            s[j] += dxh;
            q[j] = fmaf(dxh, xh, q[j]);
          }
        }
      }
    }
  }
  gng_block_reduce(s, q, C, col, pl, PP, active, red_s, red_q, gs, gq);
  if (threadIdx.x < 32) {
    float* o = partials + (((int64_t)n * Gn + chunk) * 32 + threadIdx.x) * 2;
    o[0] = gs[threadIdx.x];
    o[1] = gq[threadIdx.x];
  }
  grid_barrier(bar, gridDim.x);
  gng_fold(partials + (int64_t)n * Gn * 64, Gn, fa, fb);
  if (threadIdx.x < 32) {
    const double m = (double)cpg * (double)HW;
    s_m1[threadIdx.x] = (float)(fa[threadIdx.x] / m);  // mean(dxhat)
    s_m2[threadIdx.x] = (float)(fb[threadIdx.x] / m);  // mean(dxhat * xhat)
  }
  __syncthreads();
  if (!active) return;
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (col * 8 + j) / cpg;
    m1[j] = s_m1[g];
    m2[j] = s_m2[g];
  }
  __half* ob = dx + (int64_t)n * HW * ld_dx + col * 8;
  for (int p = p0 + pl; p < p1; p += kGngUnrollB * PP) {
    half8 rx[kGngUnrollB], rd[kGngUnrollB], ro[kGngUnrollB];
#pragma unroll
    for (int u = 0; u < kGngUnrollB; ++u)
      if (p + u * PP < p1) {
        rx[u] = ld8(xb + (int64_t)(p + u * PP) * ldx);
        rd[u] = ld8(db + (int64_t)(p + u * PP) * ld_dy);
        if (accumulate) ro[u] = ld8(ob + (int64_t)(p + u * PP) * ld_dx);
      }
#pragma unroll
    for (int u = 0; u < kGngUnrollB; ++u) {
      if (p + u * PP < p1) {
        float v[8], d[8], o[8];
        unpack8(rx[u], v);
        unpack8(rd[u], d);
        if (accumulate) unpack8(ro[u], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dv = d[j];
          if (silu) dv *= silu_grad_f(fmaf(v[j], rs[j] * G[j], Bc[j]));
          const float dxh = dv * G[j];
          const float xh = (v[j] - mu[j]) * rs[j];
          const float r = rs[j] * (dxh - m1[j] - xh * m2[j]);
          o[j] = accumulate ? o[j] + r : r;
        }
        st8(ob + (int64_t)(p + u * PP) * ld_dx, pack8(o));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
static int gng_check(const char* what, int64_t N, int64_t HW, int64_t C, int64_t Gn) {
  CGD_CHECK_ARG(N > 0 && HW > 0, "%s: bad dims", what);
  CGD_CHECK_ARG(C >= 64 && C % 64 == 0 && C <= 2048, "%s: C=%lld must be a multiple of 64 in [64, 2048]", what, (long long)C);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    CGD_CUDA(cudaGetDevice(&dev));
    CGD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // the grid barrier needs every CTA resident: one 512-thread CTA per SM
  CGD_CHECK_ARG(Gn >= 1 && N * Gn <= sms, "%s: %lld x %lld CTAs exceed the %d SMs (grid barrier needs a resident grid)", what, (long long)N,
                (long long)Gn, sms);
  return 0;
}

int launch_gn_fwd_grid(const CgdOp& op, cudaStream_t st) {
  const int64_t N = op.i[0], HW = op.i[1], C = op.i[2], ldx = op.i[3], ldy = op.i[4], Gn = op.i[5];
  if (int rc = gng_check("gn_fwd_grid", N, HW, C, Gn)) return rc;
  CGD_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && op.p[0] && op.p[1] && op.p[2] && op.p[4] && op.p[5] && op.p[6] && op.p[7], "gn_fwd_grid: bad args");
  CGD_CUDA(launch_pdl(gn_fwd_grid_kernel, dim3((unsigned)(N * Gn)), dim3(kGngThreads), 0, st, (const __half*)op.p[0], (const float*)op.p[1],
                      (const float*)op.p[2], (const float*)op.p[3], (__half*)op.p[4], (float*)op.p[5], (float*)op.p[6], (unsigned int*)op.p[7],
                      (int)HW, (int)C, ldx, ldy, (int)Gn, op.f[0], (int)(op.flags & 1)));
  return 0;
}

int launch_gn_bwd_grid(const CgdOp& op, cudaStream_t st) {
  const int64_t N = op.i[0], HW = op.i[1], C = op.i[2], ld_dy = op.i[3], ldx = op.i[4], ld_dx = op.i[5], Gn = op.i[6];
  if (int rc = gng_check("gn_bwd_grid", N, HW, C, Gn)) return rc;
  CGD_CHECK_ARG(ld_dy % 8 == 0 && ldx % 8 == 0 && ld_dx % 8 == 0 && op.p[0] && op.p[1] && op.p[2] && op.p[3] && op.p[4] && op.p[6] && op.p[7] &&
                    op.p[8],
                "gn_bwd_grid: bad args");
  CGD_CUDA(launch_pdl(gn_bwd_grid_kernel, dim3((unsigned)(N * Gn)), dim3(kGngThreads), 0, st, (const __half*)op.p[0], (const __half*)op.p[1],
                      (const float*)op.p[2], (const float*)op.p[3], (const float*)op.p[4], (const float*)op.p[5], (__half*)op.p[6],
                      (float*)op.p[7], (unsigned int*)op.p[8], (int)HW, (int)C, ld_dy, ldx, ld_dx, (int)Gn, (int)(op.flags & 1),
                      (int)((op.flags & 2) ? 1 : 0)));
  return 0;
}

}  // namespace cgd
