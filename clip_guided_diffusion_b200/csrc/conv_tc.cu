// conv_tc.cu -- implicit-GEMM convolution / GEMM on Blackwell tcgen05 tensor cores.
//
// Replaces (forward and input-gradient directions) the [3P] guided-diffusion UNet Conv2d 3x3 / 1x1 and
// Conv1d k=1 layers and the [3P] CLIP ViT Linear / patch-conv layers that the reference runs through
// cuDNN / cuBLAS.
//
// Formulation.  Activations are pixel-major fp16 ([n, y, x, c], c contiguous).  One CTA owns a
// 128-pixel x BN-channel output tile; the 128 pixels are a TW x TH x TN box of the image grid.
// For every filter tap and every 64-channel slice of the input, one TMA 4-D tiled load fetches the
// (dx,dy)-shifted box -- out-of-image coordinates are zero-filled by the TMA unit, which *is* the
// conv padding -- into a 128B-swizzled K-major shared-memory tile; one 2-D TMA load fetches the
// matching [BN x 64] slice of the pre-packed weights.  A single thread issues
// tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per stage, accumulating in TMEM.
// Pipeline: STAGES-deep smem ring with full/empty mbarriers (TMA producer warp <-> MMA warp),
// tcgen05.commit releases stages and finally signals the 4 epilogue warps, which read the
// accumulator with tcgen05.ld (32 lanes x 32 columns), add bias / residual and store fp16.
// Small-M layers (8x8, 16x16 feature maps with K up to 18k) use split-K over gridDim.z with an fp32
// workspace and a reduce kernel, so that the weight stream is spread over many SMs.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "conv_splitk.cuh"
#include "conv_sched.cuh"
#include "conv_tc.cuh"
#include "pdl.cuh"
#include "tc_ptx.cuh"

namespace cgd {

// ---------------------------------------------------------------- kernel
constexpr int BM = 128, BK = 64;
constexpr int kThreads = 192;  // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue

template <int BN>
struct TcCfg {
  static constexpr int kStages = (BN >= 192) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int mt = blockIdx.x;
  const int tw_i = mt % p.tiles_w;
  const int th_i = (mt / p.tiles_w) % p.tiles_h;
  const int tn_i = mt / (p.tiles_w * p.tiles_h);
  const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
  const int ncol0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      const int cblks = p.Cin / BK;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        const int tap = kb / cblks, cb = kb - tap * cblks;
        int dy = 0, dx = 0;
        if (p.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
        tma_load_4d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], cb * BK, w0 + dx, h0 + dy, n0);
        tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK, ncol0);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
        const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (>>4) address field
          tc_mma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue warps: TMEM lane quadrant = warp % 4
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // accumulator row = pixel within the tile
    const int w_off = r % p.TW, h_off = (r / p.TW) % p.TH, n_off = r / (p.TW * p.TH);
    const int n = n0 + n_off, h = h0 + h_off, w = w0 + w_off;
    const bool row_ok = (n < p.NB) && (h < p.H) && (w < p.W);
    if (p.pf_bytes > 0)
      l2_prefetch_slice(p.pf_ptr, p.pf_bytes, (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)),
                        (int)(gridDim.x * gridDim.y * gridDim.z), (int)threadIdx.x - 64, 128);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    if (p.splits > 1) {
      // raw fp32 partials -> workspace [split][tile row][Npad]
      float* ws = p.ws + ((size_t)blockIdx.z * p.ws_rows + (size_t)mt * BM + r) * p.Npad + ncol0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t v[16];
        __syncwarp();
        tc_ld_32x16(taddr_row + c, v);
        tc_ld_wait();
        if (row_ok)  // rows outside the image (partial tiles at the 8x8 / 16x16 levels) are never read back
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(ws + c + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      if (p.fuse_reduce) splitk_fused_reduce<BN>(p, mt, (int)blockIdx.y, (int)gridDim.y, (int)blockIdx.z, r, n, h, w, row_ok, threadIdx.x == 64);
    } else {
      const int64_t o_off = (int64_t)n * p.out_sn + (int64_t)h * p.out_sh + (int64_t)w * p.out_sw;
      const int64_t r_off = (int64_t)n * p.res_sn + (int64_t)h * p.res_sh + (int64_t)w * p.res_sw;
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t v[16];
        __syncwarp();                   // reconverge: the TMEM load is warp-collective (.sync.aligned)
        tc_ld_32x16(taddr_row + c, v);
        tc_ld_wait();
        const int col = ncol0 + c;
        if (row_ok && col < p.Cout) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
        if (col + 16 <= p.Cout && !p.out_f32 && p.out_sc == 1) {
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + col + j);
              acc[j] += b.x; acc[j + 1] += b.y; acc[j + 2] += b.z; acc[j + 3] += b.w;
            }
          }
          if (p.res) {
            float rr[16];
            unpack8(ld8(p.res + r_off + col), rr);
            unpack8(ld8(p.res + r_off + col + 8), rr + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += rr[j];
          }
          __half* o = reinterpret_cast<__half*>(p.out) + o_off + col;
          st8(o, pack8(acc));
          st8(o + 8, pack8(acc + 8));
        } else {
          for (int j = 0; j < 16 && col + j < p.Cout; ++j) {
            float a = acc[j];
            if (p.bias) a += p.bias[col + j];
            if (p.res) a += __half2float(p.res[r_off + col + j]);
            if (p.out_f32) reinterpret_cast<float*>(p.out)[o_off + (col + j) * p.out_sc] = a;
            else reinterpret_cast<__half*>(p.out)[o_off + (col + j) * p.out_sc] = __float2half_rn(a);
          }
        }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// split-K second pass: sum partials, add bias / residual, store.  One thread per 4 consecutive columns (float4 partial
// loads, splits unrolled by 4 for memory-level parallelism).
__global__ void conv_splitk_reduce_kernel(const ConvTcParams p, int m_tiles) {
  pdl_wait();
  pdl_launch_dependents();
  const int cq = (p.Cout + 3) / 4;
  const int64_t total = (int64_t)m_tiles * BM * cq;
  const size_t split_stride = (size_t)p.ws_rows * p.Npad;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(idx % cq) * 4;
    const int64_t trow = idx / cq;
    const int r = (int)(trow % BM);
    const int mt = (int)(trow / BM);
    const int tw_i = mt % p.tiles_w, th_i = (mt / p.tiles_w) % p.tiles_h, tn_i = mt / (p.tiles_w * p.tiles_h);
    const int w = tw_i * p.TW + r % p.TW, h = th_i * p.TH + (r / p.TW) % p.TH, n = tn_i * p.TN + r / (p.TW * p.TH);
    if (n >= p.NB || h >= p.H || w >= p.W) continue;
    const float* src = p.ws + (size_t)trow * p.Npad + col;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 4 <= p.splits; s += 4) {
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)(s + 0) * split_stride));
      const float4 v1 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)(s + 1) * split_stride));
      const float4 v2 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)(s + 2) * split_stride));
      const float4 v3 = __ldcs(reinterpret_cast<const float4*>(src + (size_t)(s + 3) * split_stride));
      a.x += (v0.x + v1.x) + (v2.x + v3.x); a.y += (v0.y + v1.y) + (v2.y + v3.y);
      a.z += (v0.z + v1.z) + (v2.z + v3.z); a.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; s < p.splits; ++s) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (size_t)s * split_stride));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    float acc[4] = {a.x, a.y, a.z, a.w};
    const int64_t o = (int64_t)n * p.out_sn + (int64_t)h * p.out_sh + (int64_t)w * p.out_sw;
    const int64_t ro = (int64_t)n * p.res_sn + (int64_t)h * p.res_sh + (int64_t)w * p.res_sw;
    if (col + 4 <= p.Cout && !p.out_f32 && p.out_sc == 1) {
      if (p.bias) {
        const float4 bv = *reinterpret_cast<const float4*>(p.bias + col);
        acc[0] += bv.x; acc[1] += bv.y; acc[2] += bv.z; acc[3] += bv.w;
      }
      if (p.res) {
        const __half2* rp = reinterpret_cast<const __half2*>(p.res + ro + col);
        const float2 r0 = __half22float2(rp[0]), r1 = __half22float2(rp[1]);
        acc[0] += r0.x; acc[1] += r0.y; acc[2] += r1.x; acc[3] += r1.y;
      }
      __half2* op = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p.out) + o + col);
      op[0] = __floats2half2_rn(acc[0], acc[1]);
      op[1] = __floats2half2_rn(acc[2], acc[3]);
    } else {
      for (int j = 0; j < 4 && col + j < p.Cout; ++j) {
        float v = acc[j];
        if (p.bias) v += p.bias[col + j];
        if (p.res) v += __half2float(p.res[ro + col + j]);
        if (p.out_f32) reinterpret_cast<float*>(p.out)[o + (col + j) * p.out_sc] = v;
        else reinterpret_cast<__half*>(p.out)[o + (col + j) * p.out_sc] = __float2half_rn(v);
      }
    }
  }
}

// ---------------------------------------------------------------- SIMT verification kernel
// Same contract, CUDA cores only, one thread per output element.  Exists so that GPU tests can
// cross-check the tcgen05 path layer by layer and so a tcgen05 regression cannot block bring-up of the
// rest of the step; never selected by the shipped plans (impl = 0).
__global__ void conv_simt_kernel(const __half* __restrict__ A, const __half* __restrict__ Wp, const ConvTcParams p,
                                 int64_t a_sn, int64_t a_sh, int64_t a_sw, int64_t ldb, int64_t b_sh, int64_t b_sn) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = (int64_t)p.NB * p.H * p.W * p.Cout;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx % p.Cout);
    int64_t pix = idx / p.Cout;
    const int w = (int)(pix % p.W);
    pix /= p.W;
    const int h = (int)(pix % p.H);
    const int n = (int)(pix / p.H);
    float acc = 0.f;
    for (int tap = 0; tap < p.taps; ++tap) {
      const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
      const int yy = h + dy, xx = w + dx;
      if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
      const __half2* a = reinterpret_cast<const __half2*>(A + n * a_sn + yy * a_sh + xx * a_sw);
      const __half2* wr = reinterpret_cast<const __half2*>(Wp + (int64_t)co * ldb + (int64_t)tap * p.Cin + (int64_t)h * b_sh + (int64_t)n * b_sn);
      for (int c = 0; c < p.Cin / 2; ++c) {
        const float2 av = __half22float2(a[c]), wv = __half22float2(wr[c]);
        acc = fmaf(av.x, wv.x, acc);
        acc = fmaf(av.y, wv.y, acc);
      }
    }
    if (p.bias) acc += p.bias[co];
    if (p.res) acc += __half2float(p.res[(int64_t)n * p.res_sn + (int64_t)h * p.res_sh + (int64_t)w * p.res_sw + co]);
    const int64_t o = (int64_t)n * p.out_sn + (int64_t)h * p.out_sh + (int64_t)w * p.out_sw + co * p.out_sc;
    if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = acc;
    else reinterpret_cast<__half*>(p.out)[o] = __float2half_rn(acc);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

static int pick_tile(int64_t W, int64_t H, int& TW, int& TH, int& TN) {
  // TW = smallest power of two >= min(W, 128); TH = largest power of two with TW*TH <= 128 not exceeding
  // the power-of-two ceiling of H; TN fills the remaining rows of the 128-pixel tile.
  int tw = 1;
  while (tw < W && tw < 128) tw <<= 1;
  int hceil = 1;
  while (hceil < H) hceil <<= 1;
  int th = 128 / tw;
  if (th > hceil) th = hceil;
  TW = tw;
  TH = th;
  TN = 128 / (tw * th);
  return 0;
}

int conv_tc_prepare(const CgdOp& op, ConvTcLaunch& L) {
  ConvTcParams& p = L.p;
  const int64_t NB = op.i[0], H = op.i[1], W = op.i[2], Cin = op.i[3], Cout = op.i[4], Npad = op.i[5], taps = op.i[6];
  const int64_t BN = op.i[16], splits = op.i[17] > 0 ? op.i[17] : 1;
  CGD_CHECK_ARG(NB > 0 && H > 0 && W > 0, "conv: bad image dims %lld %lld %lld", (long long)NB, (long long)H, (long long)W);
  // tile counts, pixel indices and the TMA box arithmetic below are 32-bit: reject sizes whose pixel count does not fit instead of
  // overflowing (a 2^31-row "image" used to spin in pick_tile's power-of-two search)
  CGD_CHECK_ARG(NB <= (1 << 24) && H <= (1 << 24) && W <= (1 << 24) && NB * H <= (int64_t)0x7fffffff && NB * H * W <= (int64_t)0x7fffffff && Cin <= (1 << 24) &&
                    Cout <= (1 << 24) && Npad <= (1 << 24),
                "conv: dims out of range (N %lld H %lld W %lld Cin %lld Cout %lld Npad %lld; N*H*W must fit 31 bits)", (long long)NB,
                (long long)H, (long long)W, (long long)Cin, (long long)Cout, (long long)Npad);
  CGD_CHECK_ARG(Cin > 0 && Cin % 64 == 0, "conv: Cin=%lld must be a positive multiple of 64", (long long)Cin);
  CGD_CHECK_ARG(taps == 1 || taps == 9, "conv: taps must be 1 or 9");
  CGD_CHECK_ARG(BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 192 || BN == 256, "conv: unsupported BN=%lld", (long long)BN);
  CGD_CHECK_ARG(Npad % BN == 0 && Cout <= Npad && Cout > 0, "conv: Npad=%lld must be a multiple of BN=%lld and >= Cout=%lld",
                (long long)Npad, (long long)BN, (long long)Cout);
  CGD_CHECK_ARG(op.p[0] && op.p[1] && op.p[4], "conv: null A / W / out pointer");
  CGD_CHECK_ARG(((uintptr_t)op.p[0] % 16) == 0 && ((uintptr_t)op.p[1] % 16) == 0 && ((uintptr_t)op.p[4] % 16) == 0,
                "conv: A / W / out must be 16-byte aligned");
  for (int k = 7; k <= 9; ++k) CGD_CHECK_ARG(op.i[k] % 8 == 0, "conv: A strides must be multiples of 8 elements (16 B)");
  p.NB = (int)NB; p.H = (int)H; p.W = (int)W; p.Cin = (int)Cin; p.Cout = (int)Cout; p.Npad = (int)Npad; p.taps = (int)taps;
  pick_tile(W, H, p.TW, p.TH, p.TN);
  p.tiles_w = (int)ceil_div(W, p.TW);
  p.tiles_h = (int)ceil_div(H, p.TH);
  p.tiles_n = (int)ceil_div(NB, p.TN);
  p.kblocks = (int)(taps * Cin / 64);
  p.splits = (int)splits;
  CGD_CHECK_ARG(p.splits >= 1 && p.splits <= p.kblocks, "conv: splits=%d out of range (kblocks=%d)", p.splits, p.kblocks);
  p.kb_per_split = (int)ceil_div(p.kblocks, p.splits);
  p.splits = (int)ceil_div(p.kblocks, p.kb_per_split);  // no empty splits
  p.out_sn = op.i[10]; p.out_sh = op.i[11]; p.out_sw = op.i[12];
  p.res_sn = op.i[13]; p.res_sh = op.i[14]; p.res_sw = op.i[15];
  p.bias = reinterpret_cast<const float*>(op.p[2]);
  p.res = reinterpret_cast<const __half*>(op.p[3]);
  p.out = op.p[4];
  p.ws = reinterpret_cast<float*>(op.p[5]);
  p.out_f32 = (op.flags & 1) ? 1 : 0;
  p.out_sc = op.i[19] > 0 ? op.i[19] : 1;
  if (!p.out_f32 && p.out_sc == 1) CGD_CHECK_ARG(op.i[10] % 8 == 0 && op.i[11] % 8 == 0 && op.i[12] % 8 == 0, "conv: fp16 out strides must be multiples of 8");
  if (p.res) CGD_CHECK_ARG(op.i[13] % 8 == 0 && op.i[14] % 8 == 0 && op.i[15] % 8 == 0 && ((uintptr_t)p.res % 16) == 0, "conv: residual must be 16-byte aligned with strides %% 8 == 0");
  const bool want_cluster = op.i[23] == 1 && op.i[18] != 1 && op.i[18] != 2;  // plan flag; the SIMT twin / forced single-CTA kernel ignore it
  if (p.splits > 1 && !want_cluster) CGD_CHECK_ARG(p.ws != nullptr, "conv: split-K needs a workspace");
  L.cluster_split = 0;
  p.sk_bar = reinterpret_cast<unsigned int*>(op.p[6]);
  p.epi_stats = nullptr;  // set below for the pair kernel when CONV flags 2 asks for epilogue statistics
  p.tail_full = 0;        // split last wave (CGD_CONV_TAIL): decided below
  L.tail_units = 0;
  L.BN = (int)BN;
  L.impl = (int)op.i[18];
  L.A = reinterpret_cast<const __half*>(op.p[0]);
  L.Wp = reinterpret_cast<const __half*>(op.p[1]);
  L.a_sn = op.i[7]; L.a_sh = op.i[8]; L.a_sw = op.i[9];
  L.ldb = op.i[22] > 0 ? op.i[22] : taps * Cin; L.b_sh = op.i[20]; L.b_sn = op.i[21];
  L.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.ws_rows = ((L.m_tiles + 1) / 2) * 2 * BM;
  L.n_tiles = (int)(Npad / BN);
  if (L.impl == 1) return 0;

  PFN_encodeTiled enc = get_encode_fn();
  CGD_CHECK_ARG(enc != nullptr, "conv: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t strides[3] = {(cuuint64_t)op.i[9] * 2, (cuuint64_t)op.i[8] * 2, (cuuint64_t)op.i[7] * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&L.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, op.p[0], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CGD_CHECK_ARG(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(A) failed with %d (dims %lld,%lld,%lld,%lld strides %lld,%lld,%lld box %d,%d,%d)",
                  (int)r, (long long)Cin, (long long)W, (long long)H, (long long)NB, (long long)strides[0], (long long)strides[1],
                  (long long)strides[2], p.TW, p.TH, p.TN);
  }
  const int64_t ldb = op.i[22] > 0 ? op.i[22] : taps * Cin;
  p.b_batched = (op.i[20] != 0 || op.i[21] != 0) ? 1 : 0;
  CGD_CHECK_ARG(ldb % 8 == 0 && op.i[20] % 8 == 0 && op.i[21] % 8 == 0, "conv: B strides must be multiples of 8 elements");
  if (p.b_batched) CGD_CHECK_ARG(p.TH == 1 && p.TN == 1 && p.tiles_w % 2 == 0 && L.impl != 2 && L.m_tiles >= 2,
                                 "conv: batched-B GEMMs need W %% 256 == 0 (one (h, n) per CTA pair) and the pair kernel");
  {
    const cuuint64_t K = (cuuint64_t)(taps * Cin);
    cuuint64_t dims[2] = {K, (cuuint64_t)Npad};
    cuuint64_t strides[1] = {(cuuint64_t)ldb * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&L.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, op.p[1], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CGD_CHECK_ARG(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    // pair kernel: 4-D map (K, rows, h, n) so batched GEMMs pick their B matrix by the tile's (h, n); plain weights use (0, 0)
    const cuuint64_t row_bytes = (cuuint64_t)ldb * 2 * (cuuint64_t)Npad;
    cuuint64_t dims4[4] = {K, (cuuint64_t)Npad, (cuuint64_t)(p.b_batched ? H : 1), (cuuint64_t)(p.b_batched ? NB : 1)};
    cuuint64_t strides4[3] = {(cuuint64_t)ldb * 2, p.b_batched ? (cuuint64_t)op.i[20] * 2 : row_bytes, p.b_batched ? (cuuint64_t)op.i[21] * 2 : row_bytes};
    cuuint32_t box4[4] = {64, (cuuint32_t)(BN / 2), 1, 1};
    cuuint32_t estr4[4] = {1, 1, 1, 1};
    r = enc(&L.tmB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, op.p[1], dims4, strides4, box4, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CGD_CHECK_ARG(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(W half) failed with %d", (int)r);
  }
  if (want_cluster) {
    CGD_CHECK_ARG(conv_cluster_split_ok(L), "conv: layer flagged for cluster split-K is not eligible (BN=%d splits=%d Cout=%d f32=%d)", L.BN, p.splits,
                  p.Cout, p.out_f32);
    L.cluster_split = 1;
  }
  // split-K reduction inside the conv kernel: needs the barrier buffer and every CTA of the launch resident at once
  {
    const int64_t units = conv_use_pair_kernel(L) ? (int64_t)((L.m_tiles + 1) / 2) * L.n_tiles * p.splits : (int64_t)L.m_tiles * L.n_tiles * p.splits;
    p.fuse_reduce = (p.splits > 1 && p.sk_bar != nullptr && units <= (conv_use_pair_kernel(L) ? 74 : 148)) ? 1 : 0;
    const char* e = getenv("CGD_CONV_FUSE_REDUCE");  // opt-in: see conv_splitk.cuh
    if (!(e && e[0] == '1')) p.fuse_reduce = 0;
  }
  // pair-kernel epilogue through shared memory + TMA tensor stores: same 4-D box geometry as the A operand, so rows outside
  // the image are clipped by the TMA unit; needs fp16 channel-contiguous output and whole 64-channel chunks
  p.epi_tma = (!L.cluster_split && conv_use_pair_kernel(L) && !p.out_f32 && p.out_sc == 1 && p.splits == 1 && BN >= 64 && Cout % 64 == 0) ? 1 : 0;
  if (const char* e = getenv("CGD_CONV_EPI_TMA")) if (e[0] == '0') p.epi_tma = 0;
  // split last wave (opt-in: CGD_CONV_TAIL=1; device-validated, measured no gain -- DESIGN.md): pair kernel with the TMA-store epilogue, no split-K, tail halves fit a wave
  L.tail_units = 0;
  p.tail_full = 0;
  L.tmB4 = L.tmB2;
  {
    static int tail_on = -1;
    if (tail_on < 0) {
      const char* e = getenv("CGD_CONV_TAIL");
      tail_on = (e && e[0] == '1') ? 1 : 0;
    }
    const int total = ((L.m_tiles + 1) / 2) * L.n_tiles;
    if (tail_on && device_sm_count() == 148 && p.epi_tma && !L.cluster_split && !p.b_batched && (BN == 128 || BN == 256) && sched_tail_pays(total, 74)) {
      const cuuint64_t K = (cuuint64_t)(taps * Cin);
      const cuuint64_t row_bytes = (cuuint64_t)ldb * 2 * (cuuint64_t)Npad;
      cuuint64_t dims4[4] = {K, (cuuint64_t)Npad, 1, 1};
      cuuint64_t strides4[3] = {(cuuint64_t)ldb * 2, row_bytes, row_bytes};
      cuuint32_t box4[4] = {64, (cuuint32_t)(BN / 4), 1, 1};
      cuuint32_t estr4[4] = {1, 1, 1, 1};
      CUresult r = enc(&L.tmB4, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, op.p[1], dims4, strides4, box4, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CGD_CHECK_ARG(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(W quarter) failed with %d", (int)r);
      p.tail_full = sched_full_tiles(total, 74);
      L.tail_units = sched_total_units(total, 74);
    }
  }
  // co-resident instantiation of the pair kernel for latency-bound launches (conv_tc2.cu, CO): at most one wave of tiles, or a short
  // K loop per tile.  Ring depth so that two CTAs fit an SM's 227 KB.  CGD_CONV_CO = 0 | 1 (default off until measured: DESIGN.md).
  p.nstages = 0;
  L.co_resident = 0;
  {
    static int co_on = -1;
    if (co_on < 0) {
      const char* e = getenv("CGD_CONV_CO");
      co_on = (e && e[0] == '1') ? 1 : 0;
    }
    const int64_t units = (int64_t)((L.m_tiles + 1) / 2) * L.n_tiles * p.splits;
    const int kb_tile = p.kb_per_split > 0 ? p.kb_per_split : p.kblocks;
    if (co_on && !L.cluster_split && conv_use_pair_kernel(L) && L.tail_units == 0 && !(op.flags & 2) && (units <= 74 || kb_tile <= 16)) {
      const int stage_bytes = 128 * 64 * 2 + (BN / 2) * 64 * 2;
      const int fixed = (BN >= 64 ? 2 * 128 * 64 * 2 : 0) + 1024 + 512 + 1024 /* per-CTA reservation */;
      int ns = ((227 * 1024) / 2 - fixed) / stage_bytes;
      ns = std::min(ns, 6);
      if (ns >= 2) {
        p.nstages = ns;
        L.co_resident = 1;
      }
    }
  }
  // flags 4: the "residual" operand is a QuickGELU pre-activation u and the epilogue computes acc * QuickGELU'(u) instead of acc + res
  p.res_mode = 0;
  if (op.flags & 4) {
    CGD_CHECK_ARG(p.epi_tma && p.res != nullptr && p.bias == nullptr, "conv: flags 4 (QuickGELU' epilogue) needs the pair kernel's TMA-store epilogue, a res "
                  "operand and no bias (BN=%d splits=%d Cout=%lld)", BN, p.splits, (long long)Cout);
    p.res_mode = 1;
  }
  // flags 2: the epilogue also reduces its output tile to per-octet sums for the GroupNorm that follows (GN_APPLY_EPI): needs the
  // TMA-store epilogue and tiles that are full and are 128 consecutive pixels of one image
  p.epi_stats = nullptr;
  if (op.flags & 2) {
    CGD_CHECK_ARG(p.epi_tma && op.p[7] != nullptr && W % p.TW == 0 && H % p.TH == 0 && p.TN == 1 && (p.TW == W || p.TH == 1) && L.m_tiles % 2 == 0,
                  "conv: epilogue statistics need the TMA-store epilogue and full 128-pixel tiles inside one image (W=%lld H=%lld tile %dx%dx%d)",
                  (long long)W, (long long)H, p.TW, p.TH, p.TN);
    p.epi_stats = reinterpret_cast<float*>(op.p[7]);
  }
  if (p.epi_tma) {
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    cuuint64_t ostr[3] = {(cuuint64_t)p.out_sw * 2, (cuuint64_t)p.out_sh * 2, (cuuint64_t)p.out_sn * 2};
    CUresult r = enc(&L.tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, p.out, dims, ostr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CGD_CHECK_ARG(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
    if (p.res) {
      cuuint64_t rstr[3] = {(cuuint64_t)p.res_sw * 2, (cuuint64_t)p.res_sh * 2, (cuuint64_t)p.res_sn * 2};
      r = enc(&L.tmRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)p.res, dims, rstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CGD_CHECK_ARG(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(res) failed with %d", (int)r);
    } else {
      L.tmRes = L.tmOut;
    }
  } else {
    L.tmOut = L.tmA;  // unused, but kernel parameters must be valid descriptors
    L.tmRes = L.tmA;
  }
  // <= 8 output channels, fp32 output, 3x3 (UNet head, stem dgrad): halo-tile mma.sync kernel (conv_narrow.cu); CGD_CONV_NARROW=0 keeps
  // the tcgen05 N = 16 tile for A/B runs
  {
    static int narrow_on = -1;
    if (narrow_on < 0) {
      const char* e = getenv("CGD_CONV_NARROW");
      narrow_on = (e && e[0] == '0') ? 0 : 1;
    }
    L.narrow = (narrow_on && conv_narrow_eligible(L)) ? 1 : 0;
  }
  // 8 x 8 images (deepest UNet level): one launch, a CTA per 8 output channels streaming its whole weight slab (conv_narrow.cu,
  // conv_small_kernel) instead of split-K tiles + a reduce launch.  Opt-in (CGD_CONV_SMALL=1): device-validated, but measured
  // break-even per launch (17.7 vs 18.6 us for the 1024 -> 1024 3x3) and -0.5 % in the step (profiles/r02_call_s_small_v3.log).
  {
    static int small_on = -1;
    if (small_on < 0) {
      const char* e = getenv("CGD_CONV_SMALL");
      small_on = (e && e[0] == '1') ? 1 : 0;
    }
    L.small = (small_on && !L.narrow && conv_small_eligible(L)) ? 1 : 0;
    L.tmS = L.tmA;
    if (L.small) {  // one TMA box per 64-channel slice: the 8 x 8 image with its zero border (out-of-image coordinates are zero-filled)
      const cuuint32_t e = taps == 9 ? 10 : 8;
      cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
      cuuint64_t strides[3] = {(cuuint64_t)op.i[9] * 2, (cuuint64_t)op.i[8] * 2, (cuuint64_t)op.i[7] * 2};
      cuuint32_t box[4] = {64, e, e, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&L.tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, op.p[0], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CGD_CHECK_ARG(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(8x8 image box) failed with %d", (int)r);
    }
  }
  return 0;
}

template <int BN>
static int launch_tc(const ConvTcLaunch& L, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    CGD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.mark();
  }
  dim3 grid(L.m_tiles, L.n_tiles, L.p.splits);
  CGD_CUDA(launch_pdl(conv_tc_kernel<BN>, grid, dim3(kThreads), Cfg::kSmemBytes, st, L.tmA, L.tmB, L.p));
  return 0;
}

// impl: 0 = auto (CTA-pair persistent kernel when the layer has at least two 128-pixel tiles, else the single-CTA kernel),
// 1 = SIMT verification twin, 2 = force single-CTA kernel, 3 = force CTA-pair kernel
bool conv_use_pair_kernel(const ConvTcLaunch& L) { return L.impl == 3 || (L.impl == 0 && L.m_tiles >= 2); }

int conv_tc_launch(const ConvTcLaunch& L, cudaStream_t st) {
  if (L.impl == 1) {
    const int64_t total = (int64_t)L.p.NB * L.p.H * L.p.W * L.p.Cout;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
    ConvTcParams p = L.p;
    p.splits = 1;
    CGD_CUDA(launch_pdl(conv_simt_kernel, dim3(blocks), dim3(256), 0, st, L.A, L.Wp, p, L.a_sn, L.a_sh, L.a_sw, L.ldb, L.b_sh, L.b_sn));
    return 0;
  }
  if (L.narrow) return conv_narrow_launch(L, st);
  if (L.small) return conv_small_launch(L, st);
  if (L.cluster_split) return conv_tc3_launch(L, st);
  int rc = 0;
  if (conv_use_pair_kernel(L)) rc = conv_tc2_launch(L, st);
  else if (L.p.b_batched) { set_error("conv: batched-B GEMM reached the single-CTA kernel"); return -1; }
  else
  switch (L.BN) {
    case 16: rc = launch_tc<16>(L, st); break;
    case 32: rc = launch_tc<32>(L, st); break;
    case 64: rc = launch_tc<64>(L, st); break;
    case 128: rc = launch_tc<128>(L, st); break;
    case 192: rc = launch_tc<192>(L, st); break;
    case 256: rc = launch_tc<256>(L, st); break;
    default: set_error("conv: unsupported BN %d", L.BN); return -1;
  }
  if (rc) return rc;
  if (L.p.splits > 1 && !L.p.fuse_reduce) {
    const int64_t total = (int64_t)L.m_tiles * BM * ((L.p.Cout + 3) / 4);
    const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 148 * 8);
    CGD_CUDA(launch_pdl(conv_splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, L.p, L.m_tiles));
  }
  return 0;
}

int conv_tc_num_launches(const ConvTcLaunch& L) {
  return (L.impl != 1 && !L.small && !L.narrow && L.p.splits > 1 && !L.p.fuse_reduce && !L.cluster_split) ? 2 : 1;
}

}  // namespace cgd
