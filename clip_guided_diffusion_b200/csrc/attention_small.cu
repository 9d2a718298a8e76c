// attention_small.cu -- single-tile softmax attention (T <= 64 keys, head dim 64), forward and backward, one CTA per
// (image, head), all five / seven products on warp-level tensor-core MMAs (mma.sync m16n8k16, fp16 operands, fp32 accumulate).
//
// Replaces [3P] CLIP nn.MultiheadAttention's core for ViT-B/32 (50 tokens) and [3P] guided-diffusion QKVAttention(Legacy) at the
// 8x8 UNet level (64 tokens).  Why a separate kernel: these problems are one 64x64 tile per head -- the
// tcgen05 pair kernel wants 256-row tiles (conv_tc2.cu) and the fp32 CUDA-core flash kernels of attention.cu spend 17 us
// (forward) and 62 us (backward, three launches) on 192 such tiles (profiles/r01_launches_cfg2_step_v5_warm.csv).  Here the
// whole backward -- S, P, dP, dS, dQ, dK, dV -- is one launch; P and dS go through shared memory once (fp16, like the
// reference's fp16 attention weights) to feed the transposed products.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ops.cuh"
#include "pdl.cuh"

namespace cgd {

constexpr int AS_T = 64;        // tile edge: queries, keys, head dim
constexpr int AS_LD = AS_T + 8; // shared-memory row pitch in halfs (144 B: 16-byte aligned, ldmatrix conflict-free)

struct AttnSmallArgs {
  const __half *q, *k, *v, *dout;
  __half *out, *dq, *dk, *dv;
  float* lse;
  int B, heads, T;
  int64_t qbs, qrs, qhs;  // qkv batch / row / head strides (elements)
  int64_t obs, ors, ohs;  // out / dout strides
  float scale;
};

__device__ __forceinline__ uint32_t as_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, fp16 row) * B(16x8, fp16 col)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// 64 x 64 fp16 tile (rows < T valid, zero-filled beyond) global -> shared, 16-byte vectors
__device__ __forceinline__ void as_load_tile(__half* s, const __half* g, int64_t rs, int T) {
  for (int v = threadIdx.x; v < AS_T * 8; v += blockDim.x) {
    const int r = v >> 3, c = (v & 7) * 8;
    half8 val;
    if (r < T) val = ld8(g + (int64_t)r * rs + c);
    else val.a = val.b = val.c = val.d = __floats2half2_rn(0.f, 0.f);
    *reinterpret_cast<half8*>(s + r * AS_LD + c) = val;
  }
}

// Fragment address helpers (lane -> row address for ldmatrix.x4); see the comment block in each product below.
// A operand 16x16 at (r0, c0) of a row-major [m][k] tile
__device__ __forceinline__ uint32_t as_addr_a(const __half* s, int r0, int c0, int lane) {
  const int mi = lane >> 3;
  return as_smem(s + (r0 + (lane & 7) + (mi & 1) * 8) * AS_LD + c0 + (mi >> 1) * 8);
}
// B operand (k16 x n16 = two n-tiles) from a tile stored [n][k] (k contiguous): regs {b0,b1} of n-tile 0, {b0,b1} of n-tile 1
__device__ __forceinline__ uint32_t as_addr_b_nk(const __half* s, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  return as_smem(s + (n0 + (lane & 7) + (mi >> 1) * 8) * AS_LD + k0 + (mi & 1) * 8);
}
// B operand (k16 x n16) from a tile stored [k][n] (n contiguous), with ldmatrix.trans: same register order as above
__device__ __forceinline__ uint32_t as_addr_b_kn(const __half* s, int k0, int n0, int lane) {
  const int mi = lane >> 3;
  return as_smem(s + (k0 + (lane & 7) + (mi & 1) * 8) * AS_LD + n0 + (mi >> 1) * 8);
}
// A operand 16x16 = (X^T)[m0.., k0..] from a tile X stored [k][m] (m contiguous), with ldmatrix.trans
__device__ __forceinline__ uint32_t as_addr_a_t(const __half* s, int m0, int k0, int lane) {
  const int mi = lane >> 3;
  return as_smem(s + (k0 + (lane & 7) + (mi >> 1) * 8) * AS_LD + m0 + (mi & 1) * 8);
}

// acc[j] (8 n-tiles of 8 columns) = X[r0..r0+16, :] * Y^T with Y stored [n][k] (k = head dim contiguous): S = Q K^T, dP = dO V^T
__device__ __forceinline__ void as_mm_nk(float (&acc)[8][4], const __half* X, const __half* Y, int r0, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    ldsm_x4(as_addr_a(X, r0, ks * 16, lane), a);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      uint32_t b[4];
      ldsm_x4(as_addr_b_nk(Y, jj * 16, ks * 16, lane), b);
      mma16816(acc[2 * jj], a, b[0], b[1]);
      mma16816(acc[2 * jj + 1], a, b[2], b[3]);
    }
  }
}
// acc = P[16 rows, 64] (fp16 A fragments held in registers, 4 k-steps) * Y with Y stored [k][n] (n contiguous): O = P V, dQ = dS K
__device__ __forceinline__ void as_mm_reg_kn(float (&acc)[8][4], const uint32_t (&pa)[4][4], const __half* Y, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      uint32_t b[4];
      ldsm_x4_t(as_addr_b_kn(Y, ks * 16, jj * 16, lane), b);
      mma16816(acc[2 * jj], pa[ks], b[0], b[1]);
      mma16816(acc[2 * jj + 1], pa[ks], b[2], b[3]);
    }
  }
}
// acc = (X^T)[m0..m0+16, :] * Y with X stored [k][m] and Y stored [k][n]: dV = P^T dO, dK = dS^T Q
__device__ __forceinline__ void as_mm_t_kn(float (&acc)[8][4], const __half* X, const __half* Y, int m0, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    ldsm_x4_t(as_addr_a_t(X, m0, ks * 16, lane), a);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      uint32_t b[4];
      ldsm_x4_t(as_addr_b_kn(Y, ks * 16, jj * 16, lane), b);
      mma16816(acc[2 * jj], a, b[0], b[1]);
      mma16816(acc[2 * jj + 1], a, b[2], b[3]);
    }
  }
}

// Row softmax of the warp's 16 x 64 score fragment (rows lane/4 and lane/4 + 8), keys >= T masked; returns P in `s`.
__device__ __forceinline__ void as_softmax(float (&s)[8][4], int T, float scale, int lane, float& lse0, float& lse1) {
  const int cbase = (lane & 3) * 2;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool ok = j * 8 + cbase + e < T;
      s[j][e] = ok ? s[j][e] * scale : -INFINITY;
      s[j][2 + e] = ok ? s[j][2 + e] * scale : -INFINITY;
      mx0 = fmaxf(mx0, s[j][e]);
      mx1 = fmaxf(mx1, s[j][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[j][e] = __expf(s[j][e] - mx0);  // exp(-inf) = 0 for masked keys; T >= 1 keeps mx finite
      s[j][2 + e] = __expf(s[j][2 + e] - mx1);
      l0 += s[j][e];
      l1 += s[j][2 + e];
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  lse0 = mx0 + __logf(l0);
  lse1 = mx1 + __logf(l1);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j][0] *= i0;
    s[j][1] *= i0;
    s[j][2] *= i1;
    s[j][3] *= i1;
  }
}
// fp32 C fragments (16 x 64) -> fp16 A fragments for the next product (4 k-steps of 16)
__device__ __forceinline__ void as_c_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_h2(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_h2(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_h2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_h2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}
// fp32 C fragments of rows r0 + lane/4 (+8) -> global fp16 rows (row stride rs), rows >= T skipped
__device__ __forceinline__ void as_store_c(const float (&c)[8][4], __half* g, int64_t rs, int r0, int T, int lane) {
  const int ra = r0 + (lane >> 2), rb = ra + 8, cb = (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (ra < T) *reinterpret_cast<__half2*>(g + (int64_t)ra * rs + j * 8 + cb) = __floats2half2_rn(c[j][0], c[j][1]);
    if (rb < T) *reinterpret_cast<__half2*>(g + (int64_t)rb * rs + j * 8 + cb) = __floats2half2_rn(c[j][2], c[j][3]);
  }
}

__global__ void __launch_bounds__(128) attn_small_fwd_kernel(const AttnSmallArgs a) {
  __shared__ __align__(16) __half Qs[AS_T * AS_LD], Ks[AS_T * AS_LD], Vs[AS_T * AS_LD];
  pdl_wait();
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qoff = (int64_t)b * a.qbs + (int64_t)h * a.qhs;
  as_load_tile(Qs, a.q + qoff, a.qrs, a.T);
  as_load_tile(Ks, a.k + qoff, a.qrs, a.T);
  as_load_tile(Vs, a.v + qoff, a.qrs, a.T);
  __syncthreads();
  const int r0 = warp * 16;
  if (r0 >= a.T) return;
  float s[8][4], lse0, lse1;
  as_mm_nk(s, Qs, Ks, r0, lane);
  as_softmax(s, a.T, a.scale, lane, lse0, lse1);
  if (a.lse && (lane & 3) == 0) {
    const int ra = r0 + (lane >> 2);
    float* l = a.lse + ((int64_t)b * a.heads + h) * a.T;
    if (ra < a.T) l[ra] = lse0;
    if (ra + 8 < a.T) l[ra + 8] = lse1;
  }
  uint32_t pa[4][4];
  as_c_to_a(s, pa);
  float o[8][4];
  as_mm_reg_kn(o, pa, Vs, lane);
  as_store_c(o, a.out + (int64_t)b * a.obs + (int64_t)h * a.ohs, a.ors, r0, a.T, lane);
}

__global__ void __launch_bounds__(128) attn_small_bwd_kernel(const AttnSmallArgs a) {
  extern __shared__ __align__(16) __half as_dyn[];
  __half *Qs = as_dyn, *Ks = Qs + AS_T * AS_LD, *Vs = Ks + AS_T * AS_LD, *dOs = Vs + AS_T * AS_LD, *Ps = dOs + AS_T * AS_LD, *dSs = Ps + AS_T * AS_LD;
  pdl_wait();
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qoff = (int64_t)b * a.qbs + (int64_t)h * a.qhs;
  const int64_t ooff = (int64_t)b * a.obs + (int64_t)h * a.ohs;
  as_load_tile(Qs, a.q + qoff, a.qrs, a.T);
  as_load_tile(Ks, a.k + qoff, a.qrs, a.T);
  as_load_tile(Vs, a.v + qoff, a.qrs, a.T);
  as_load_tile(dOs, a.dout + ooff, a.ors, a.T);
  __syncthreads();
  const int r0 = warp * 16;
  {
    // ---- rows r0 .. r0+15 of the queries: P, dP, dS, dQ
    float s[8][4], dp[8][4], lse0, lse1;
    as_mm_nk(s, Qs, Ks, r0, lane);
    as_softmax(s, a.T, a.scale, lane, lse0, lse1);   // s = P (rows of padded queries are harmless: their dO rows are zero)
    as_mm_nk(dp, dOs, Vs, r0, lane);     // dP = dO V^T
    float d0 = 0.f, d1 = 0.f;            // delta = rowsum(P * dP) (= rowsum(dO * O))
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
      d1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // dp <- dS = P * (dP - delta) * scale
      dp[j][0] = s[j][0] * (dp[j][0] - d0) * a.scale;
      dp[j][1] = s[j][1] * (dp[j][1] - d0) * a.scale;
      dp[j][2] = s[j][2] * (dp[j][2] - d1) * a.scale;
      dp[j][3] = s[j][3] * (dp[j][3] - d1) * a.scale;
    }
    // P and dS to shared memory (fp16, [q][key]) for the transposed products
    const int ra = r0 + (lane >> 2), cb = (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<__half2*>(Ps + ra * AS_LD + j * 8 + cb) = __floats2half2_rn(s[j][0], s[j][1]);
      *reinterpret_cast<__half2*>(Ps + (ra + 8) * AS_LD + j * 8 + cb) = __floats2half2_rn(s[j][2], s[j][3]);
      *reinterpret_cast<__half2*>(dSs + ra * AS_LD + j * 8 + cb) = __floats2half2_rn(dp[j][0], dp[j][1]);
      *reinterpret_cast<__half2*>(dSs + (ra + 8) * AS_LD + j * 8 + cb) = __floats2half2_rn(dp[j][2], dp[j][3]);
    }
    uint32_t dsa[4][4];
    as_c_to_a(dp, dsa);
    float dq[8][4];
    as_mm_reg_kn(dq, dsa, Ks, lane);     // dQ = dS K
    as_store_c(dq, a.dq + qoff, a.qrs, r0, a.T, lane);
  }
  __syncthreads();
  {
    // ---- rows r0 .. r0+15 of the keys: dV = P^T dO, dK = dS^T Q
    float acc[8][4];
    as_mm_t_kn(acc, Ps, dOs, r0, lane);
    as_store_c(acc, a.dv + qoff, a.qrs, r0, a.T, lane);
    as_mm_t_kn(acc, dSs, Qs, r0, lane);
    as_store_c(acc, a.dk + qoff, a.qrs, r0, a.T, lane);
  }
}

int attn_small_supported(int64_t T) { return T >= 1 && T <= AS_T; }

static int attn_small_args(const CgdOp& op, AttnSmallArgs& a, bool bwd) {
  a.B = (int)op.i[0]; a.heads = (int)op.i[1]; a.T = (int)op.i[2];
  a.qbs = op.i[4]; a.qrs = op.i[5]; a.qhs = op.i[6]; a.obs = op.i[7]; a.ors = op.i[8]; a.ohs = op.i[9];
  a.scale = op.f[0];
  a.q = (const __half*)op.p[0]; a.k = (const __half*)op.p[1]; a.v = (const __half*)op.p[2];
  if (!bwd) {
    a.out = (__half*)op.p[3];
    a.lse = (float*)op.p[4];
  } else {
    a.dout = (const __half*)op.p[4];
    a.dq = (__half*)op.p[6]; a.dk = (__half*)op.p[7]; a.dv = (__half*)op.p[8];
  }
  return 0;
}

int launch_attn_small_fwd(const CgdOp& op, cudaStream_t st) {
  AttnSmallArgs a{};
  attn_small_args(op, a, false);
  CGD_CUDA(launch_pdl(attn_small_fwd_kernel, dim3(a.heads, a.B), dim3(128), 0, st, a));
  return 0;
}
int launch_attn_small_bwd(const CgdOp& op, cudaStream_t st) {
  AttnSmallArgs a{};
  attn_small_args(op, a, true);
  constexpr int smem = 6 * AS_T * AS_LD * (int)sizeof(__half);
  static bool set = false;
  if (!set) {
    CGD_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    set = true;
  }
  CGD_CUDA(launch_pdl(attn_small_bwd_kernel, dim3(a.heads, a.B), dim3(128), smem, st, a));
  return 0;
}

}  // namespace cgd
