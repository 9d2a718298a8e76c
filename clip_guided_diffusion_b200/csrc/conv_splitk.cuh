// conv_splitk.cuh -- split-K reduction fused into the conv kernels (no second launch).
//
// The 8x8 .. 32x32 UNet layers and the long-K ViT GEMMs split K over several CTAs (or CTA pairs) that each hold an fp32 partial
// tile.  Every CTA writes its partial to the L2-resident workspace, the S CTAs that share an output row tile meet at a
// generation-counted barrier in global memory (they are co-resident: the plan never splits beyond one wave), and each then
// reduces its share of the tile's 8-column chunks over all S partials in a fixed order (deterministic), adds bias / residual and
// stores.  Alternative to conv_splitk_reduce_kernel (239 launches per cfg2 step).  Opt-in (CGD_CONV_FUSE_REDUCE=1): measured
// with fence-based barriers it LOST 2 ms per step against the separate reduce kernel (65.4 vs 75.3 steps/s) -- a software barrier
// in global memory costs more than a kernel boundary on this part.
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace cgd {

// Called by the 128 epilogue threads of a CTA after they stored their rows of the partial tile.  r = accumulator row (pixel of the
// 128-pixel tile `mt`), (n, h, w) its image coordinates, `issuer` = the one thread that talks to the barrier.
template <int BN>
__device__ __forceinline__ void splitk_fused_reduce(const ConvTcParams& p, int mt, int n_tile, int n_tiles, int split, int r, int n, int h, int w,
                                                    bool row_ok, bool issuer) {
  named_bar_sync(1, 128);  // all partial rows of this CTA are stored; the issuer's release atomic publishes them (cumulative)
  if (issuer) global_barrier_arrive_wait(p.sk_bar + 2 * ((size_t)mt * n_tiles + n_tile), (unsigned)p.splits);
  named_bar_sync(1, 128);
  if (!row_ok) return;
  const int ncol0 = n_tile * BN;
  const size_t split_stride = (size_t)p.ws_rows * p.Npad;
  const float* base = p.ws + ((size_t)mt * 128 + r) * p.Npad + ncol0;
  const int64_t o_off = (int64_t)n * p.out_sn + (int64_t)h * p.out_sh + (int64_t)w * p.out_sw;
  const int64_t r_off = (int64_t)n * p.res_sn + (int64_t)h * p.res_sh + (int64_t)w * p.res_sw;
  const bool vec_ok = !p.out_f32 && p.out_sc == 1;
  for (int c = split; c < BN / 8; c += p.splits) {  // this CTA's 8-column chunks of the tile
    const int col = ncol0 + c * 8;
    if (col >= p.Cout) break;
    float a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = 0.f;
    const float* src = base + c * 8;
#pragma unroll 4
    for (int s = 0; s < p.splits; ++s) {  // fixed order: deterministic
      const float4 x0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)s * split_stride));
      const float4 x1 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)s * split_stride + 4));
      a[0] += x0.x; a[1] += x0.y; a[2] += x0.z; a[3] += x0.w;
      a[4] += x1.x; a[5] += x1.y; a[6] += x1.z; a[7] += x1.w;
    }
    if (col + 8 <= p.Cout && vec_ok) {
      if (p.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
        a[0] += b0.x; a[1] += b0.y; a[2] += b0.z; a[3] += b0.w;
        a[4] += b1.x; a[5] += b1.y; a[6] += b1.z; a[7] += b1.w;
      }
      if (p.res) {
        float rr[8];
        unpack8(ld8(p.res + r_off + col), rr);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += rr[e];
      }
      st8(reinterpret_cast<__half*>(p.out) + o_off + col, pack8(a));
    } else {
      for (int e = 0; e < 8 && col + e < p.Cout; ++e) {
        float x = a[e];
        if (p.bias) x += p.bias[col + e];
        if (p.res) x += __half2float(p.res[r_off + col + e]);
        if (p.out_f32) reinterpret_cast<float*>(p.out)[o_off + (col + e) * p.out_sc] = x;
        else reinterpret_cast<__half*>(p.out)[o_off + (col + e) * p.out_sc] = __float2half_rn(x);
      }
    }
  }
}

}  // namespace cgd
