// conv_tc3.cu -- split-K inside a thread-block cluster: the latency-bound layers (8x8 ... 32x32 UNet levels, the ViT's K = 2304 /
// 3072 Linears at M = 800) without the fp32 partial tensor and without the second (reduce) launch.
//
// What it replaces (measured, profiles/r01_launches_v11_warm.csv): those layers ran the pair kernel over (tile, K split) units,
// wrote S fp32 partial tiles to global memory and summed them in conv_splitk_reduce_kernel -- 153 reduce launches of ~5.5 us per
// step, and e.g. 24 us for a 3072 -> 768 Linear whose mainloop is ~5 us.  A software barrier among co-resident CTAs had lost
// against the kernel boundary (conv_splitk.cuh); the hardware cluster barrier does not.
//
// One cluster of 2*S CTAs (S <= 8: sixteen CTAs need the non-portable cluster size) owns ONE output tile of 256 pixels x BN
// channels.  CTA pair s (cluster ranks 2s, 2s+1) is a cta_group::2 MMA pair exactly as in conv_tc2.cu and accumulates K-blocks
// [s*kps, (s+1)*kps) into its own TMEM accumulator.  Then
//   barrier.cluster  (every pair's MMAs have retired: the operand ring of every CTA is free)
//   reduce-scatter through distributed shared memory: pair s pushes the columns owned by pair j (BN/S of them) as fp32 into
//     CTA (2j + h)'s ring memory, slot s, column-major float4 units -> 512 contiguous bytes per warp store, conflict-free reads
//   barrier.cluster  (release / acquire: the remote stores are visible)
//   pair j sums its S slots in a fixed order (bit-reproducible), adds bias / residual, converts and stores its BN/S columns.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "conv_tc.cuh"
#include "pdl.cuh"
#include "tc_ptx.cuh"

namespace cgd {

constexpr int BM3 = 128, BK3 = 64;
constexpr int kThreads3 = 192;  // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue

template <int BN>
struct Tc3Cfg {
  static constexpr int kABytes = BM3 * BK3 * 2;       // this CTA's 128 pixel rows
  static constexpr int kBBytes = (BN / 2) * BK3 * 2;  // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBudget = 227 * 1024 - 1024 - 512;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int kSlotBytesTotal = BN * BM3 * 4;  // S slots x (BN / S) columns x 128 rows x fp32
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
  static_assert(kSlotBytesTotal <= kStages * kStageBytes, "the reduction slots reuse the operand ring");
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreads3, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p, int n_tiles, int S) {
  using Cfg = Tc3Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t s = rank >> 1, h = rank & 1u, lrank = rank & ~1u;  // K split, pixel half of the pair tile, the pair's leader
  const bool leader = h == 0;
  const int tile = blockIdx.x / (2 * S);
  const int n_tile = tile % n_tiles, pair = tile / n_tiles;
  const int mt = 2 * pair + (int)h;
  const int kb0 = (int)s * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kblocks);  // non-empty for every s (host)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int k = 0; k < Cfg::kStages; ++k) {
      mbar_init(&full_bar[k], 2);   // the leader's copy is used: one arrive per CTA of the pair (+ the TMA transaction bytes)
      mbar_init(&empty_bar[k], 1);  // multicast tcgen05.commit
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  const int tw_i = mt % p.tiles_w, th_i = (mt / p.tiles_w) % p.tiles_h, tn_i = mt / (p.tiles_w * p.tiles_h);
  const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (one thread per CTA)
      const int cblks = p.Cin / BK3;
      const int bcol = n_tile * BN + (int)h * (BN / 2);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const int tap = kb / cblks, cb = kb - tap * cblks;
        int dy = 0, dx = 0;
        if (p.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
        tma2_load_4d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], cb * BK3, w0 + dx, h0 + dy, n0);
        tma2_load_4d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK3, bcol, 0, 0);
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
        else mbar_arrive_remote(&full_bar[stage], lrank);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {  // ===== MMA issuer: one thread of the pair's leader
      constexpr uint32_t idesc = make_idesc_f16_mn(256, BN);
      const uint16_t mask = (uint16_t)(0x3u << lrank);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
        const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
        for (int k = 0; k < BK3 / 16; ++k) tc2_mma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        tc2_commit_mc(&empty_bar[stage], mask);
        if (kb == kb1 - 1) tc2_commit_mc(tmem_full_bar, mask);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    mbar_wait(tmem_full_bar, 0);  // this pair's accumulator is complete (and its MMAs no longer read the ring)
    tc_fence_after();
  }
  __syncwarp();
  cluster_sync_all();  // A: every CTA's ring memory is free

  const int cw = BN / S;  // columns owned by one pair (multiple of 16, host)
  const int quad = warp & 3;
  const int r = quad * 32 + lane;
  const uint32_t slots = smem_u32(smem);
  const uint32_t slot_bytes = (uint32_t)cw * BM3 * 4;
  if (warp >= 2) {
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t my_off = s * slot_bytes + (uint32_t)r * 16;
    // 16-column units, software-pipelined (tcgen05.wait::ld waits for ALL outstanding loads, so unit u+1 is issued right after
    // the wait and its latency hides behind the four remote stores of unit u).  The units start at the columns of pair s+1 and
    // wrap around: at any moment the S senders of a half address S different receivers (DSMEM moves ~20 B/clk per SM).
    constexpr int kUnits = BN / 16;
    const int rot = (int)((s + 1) % (uint32_t)S) * (cw / 16);
    auto unit_col = [&](int u) {
      int x = u + rot;
      return (x >= kUnits ? x - kUnits : x) * 16;
    };
    auto push = [&](int col, const uint32_t* v) {
      const int j = col / cw, c = col - j * cw;
      const uint32_t dst = mapa_u32(slots + my_off, (uint32_t)(2 * j) + h) + (uint32_t)((c / 4) * BM3 * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) st_cluster_v4(dst + (uint32_t)(q * BM3 * 16), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    };
    uint32_t va[16], vb[16];
    __syncwarp();
    tc_ld_32x16(taddr_row + (uint32_t)unit_col(0), va);
    tc_ld_wait();
#pragma unroll 1
    for (int u = 0; u < kUnits; u += 2) {  // kUnits is even for every BN
      tc_ld_32x16(taddr_row + (uint32_t)unit_col(u + 1), vb);
      push(unit_col(u), va);
      tc_ld_wait();
      if (u + 2 < kUnits) tc_ld_32x16(taddr_row + (uint32_t)unit_col(u + 2), va);
      push(unit_col(u + 1), vb);
      tc_ld_wait();
    }
    tc_fence_before();
  }
  __syncwarp();
  cluster_sync_all();  // B: all partial sums have landed

  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
  if (warp >= 2) {
    const int w = w0 + r % p.TW, hh = h0 + (r / p.TW) % p.TH, n = n0 + r / (p.TW * p.TH);
    const bool row_ok = (n < p.NB) && (hh < p.H) && (w < p.W);
    const int col0 = n_tile * BN + (int)s * cw;
    const int64_t o_off = (int64_t)n * p.out_sn + (int64_t)hh * p.out_sh + (int64_t)w * p.out_sw;
    const int64_t r_off = (int64_t)n * p.res_sn + (int64_t)hh * p.res_sh + (int64_t)w * p.res_sw;
    const uint8_t* mine = smem + (size_t)r * 16;
    if (row_ok) {
#pragma unroll 1
      for (int c = 0; c < cw; c += 8) {
        const int col = col0 + c;
        if (col >= p.Cout) break;  // Cout % 8 == 0 (host): whole 8-column units only
        float a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = 0.f;
        for (int j = 0; j < S; ++j) {  // fixed order: bit-reproducible
          const float4 u0 = *reinterpret_cast<const float4*>(mine + (size_t)j * slot_bytes + (size_t)(c / 4) * BM3 * 16);
          const float4 u1 = *reinterpret_cast<const float4*>(mine + (size_t)j * slot_bytes + (size_t)(c / 4 + 1) * BM3 * 16);
          a[0] += u0.x; a[1] += u0.y; a[2] += u0.z; a[3] += u0.w;
          a[4] += u1.x; a[5] += u1.y; a[6] += u1.z; a[7] += u1.w;
        }
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
      }
    }
  }
}

template <int BN>
static int launch_tc3(const ConvTcLaunch& L, cudaStream_t st) {
  using Cfg = Tc3Cfg<BN>;
  const int S = L.p.splits;
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    CGD_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    CGD_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<BN>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set.mark();
  }
  const int pair_tiles = (L.m_tiles + 1) / 2;
  const int tiles = pair_tiles * L.n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles * 2 * S));
  cfg.blockDim = dim3(kThreads3);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = (unsigned)(2 * S);
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attrs;
  CGD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc3_kernel<BN>, L.tmA, L.tmB2, L.p, L.n_tiles, S));
  return 0;
}

template <int BN>
static int max_clusters_tc3(int S) {
  using Cfg = Tc3Cfg<BN>;
  if (cudaFuncSetAttribute(conv_tc3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(conv_tc3_kernel<BN>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * S * 64));
  cfg.blockDim = dim3(kThreads3);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = (unsigned)(2 * S);
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_tc3_kernel<BN>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}
// co-resident clusters of 2*S CTAs the device can hold for tile width BN (cudaOccupancyMaxActiveClusters); -1 without a device
int conv_tc3_max_clusters(int BN, int S) {
  switch (BN) {
    case 64: return max_clusters_tc3<64>(S);
    case 128: return max_clusters_tc3<128>(S);
    case 192: return max_clusters_tc3<192>(S);
    case 256: return max_clusters_tc3<256>(S);
    default: return -1;
  }
}

// eligibility of a layer the plan flagged (op.i[23] = 1); S = p.splits after the no-empty-split rounding
bool conv_cluster_split_ok(const ConvTcLaunch& L) {
  const ConvTcParams& p = L.p;
  const int S = p.splits;
  return (L.impl == 0 || L.impl == 3) && !p.out_f32 && p.out_sc == 1 && !p.b_batched && p.Cout % 8 == 0 && L.BN >= 64 && S >= 2 && S <= 8 &&
         L.BN % S == 0 && (L.BN / S) % 16 == 0;
}

int conv_tc3_launch(const ConvTcLaunch& L, cudaStream_t st) {
  switch (L.BN) {
    case 64: return launch_tc3<64>(L, st);
    case 128: return launch_tc3<128>(L, st);
    case 192: return launch_tc3<192>(L, st);
    case 256: return launch_tc3<256>(L, st);
    default: set_error("conv (cluster split-K): unsupported BN %d", L.BN); return -1;
  }
}

}  // namespace cgd
