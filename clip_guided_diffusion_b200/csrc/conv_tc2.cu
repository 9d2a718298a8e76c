// conv_tc2.cu -- persistent CTA-pair variant of the implicit-GEMM conv (see conv_tc.cu for the formulation).
//
// What changes against conv_tc_kernel (measured on B200, ncu, 256x256x256->256 3x3: tensor pipe 33 % active, 3.46 waves of
// one-tile CTAs, un-overlapped epilogue, 48 KB of L2->smem traffic per K-block per CTA):
//   * cta_group::2: two CTAs of a cluster (one TPC) form a 256-pixel x BN tile.  Each CTA stages its own 128 pixel rows of
//     A and only HALF of the weight tile (BN/2 rows); tcgen05.mma.cta_group::2 (M=256) reads B from both CTAs' shared
//     memory, so the weight traffic per CTA halves (48 KB -> 32 KB per K-block at BN=256) and the ring gets 6 stages.
//   * persistent: grid = one cluster per SM pair, static round-robin over (split, pixel-pair tile, channel tile).
//   * two TMEM accumulator stages (2 x BN columns): the epilogue warps drain tile i while the MMA warp already works on
//     tile i+1 (tmem_full / tmem_empty mbarriers; tmem_empty is arrived remotely by the peer CTA's epilogue threads).
// Protocol (DeepGEMM / CUTLASS sm100 2-SM skeleton): both CTAs run a TMA producer whose loads complete_tx on the LEADER's
// full barrier (peer-bit-masked address); the leader's single MMA thread issues the MMAs and multicasts tcgen05.commit to
// the empty barriers / tmem_full barriers of both CTAs.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "conv_sched.cuh"
#include "conv_splitk.cuh"
#include "conv_tc.cuh"
#include "pdl.cuh"
#include "tc_ptx.cuh"

namespace cgd {

constexpr int BM2 = 128, BK2 = 64;
constexpr int kThreads2 = 192;  // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..5 epilogue
constexpr int kEpiThreads2 = 128;

constexpr int kEpiChunkBytes = BM2 * 64 * 2;  // one [128 rows x 64 channels] fp16 staging tile (SWIZZLE_128B)

template <int BN>
struct Tc2Cfg {
  static constexpr int kABytes = BM2 * BK2 * 2;          // 16 KB: this CTA's 128 pixel rows
  static constexpr int kBBytes = (BN / 2) * BK2 * 2;     // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue staging: 2 output tiles (TMA store sources); residual tiles are TMA-loaded INTO them and updated in place
  static constexpr int kEpiBytes = BN >= 64 ? 2 * kEpiChunkBytes : 0;
  static constexpr int kStatBytes = 512;  // STATS instantiation only: epilogue statistics exchange, 2 buffers x 4 warps x 16 floats
  static constexpr int kBudget = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - kEpiBytes;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 + 512;
  static_assert(kSmemBytes + kStatBytes <= 227 * 1024, "the STATS instantiation appends kStatBytes");
};

struct Tile2 {
  int mt, n_tile, kb0, kb1, split;
};
// tile index -> (split, pixel-pair tile, channel tile); channel tile fastest so neighbouring clusters share A in L2
__device__ __forceinline__ Tile2 decode_tile(const ConvTcParams& p, int t, int n_tiles, int pair_tiles, uint32_t rank) {
  Tile2 r;
  r.n_tile = t % n_tiles;
  const int rest = t / n_tiles;
  const int pair = rest % pair_tiles;
  r.split = rest / pair_tiles;
  r.mt = 2 * pair + (int)rank;
  r.kb0 = r.split * p.kb_per_split;
  r.kb1 = min(r.kb0 + p.kb_per_split, p.kblocks);
  return r;
}

// STATS: the epilogue additionally reduces every output chunk to per-octet sums for the GroupNorm that follows (CONV flags 2); a
// separate instantiation so that the default kernel's code and shared-memory layout are untouched
// TAIL: the tiles of the last partial wave are cut in two BN/2-wide halves (conv_sched.cuh); `total_tiles` then counts schedule units
// (p.tail_full whole tiles first, then the halves) and tmB4 loads a quarter of the weight tile per CTA.  Also a separate instantiation.
// CO: "co-resident" instantiation for latency-bound launches (one wave of tiles or a short K loop): compiled for two CTAs per SM
// (<= 168 registers) and launched with a short operand ring (p.nstages), so that with programmatic dependent launch the NEXT kernel's
// CTAs are resident -- barriers initialised, descriptors prefetched, TMEM allocated where columns are free -- while this one drains.
template <int BN, bool STATS = false, bool TAIL = false, bool CO = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, CO ? 2 : 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmB4, const ConvTcParams p, int n_tiles,
                int pair_tiles, int total_tiles) {
  using Cfg = Tc2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.nstages;  // operand-ring depth of this launch (host: 2 <= nstages <= Cfg::kStages; shared memory sized for it)
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NS * Cfg::kABytes;
  uint8_t* smem_epi = smem + NS * Cfg::kStageBytes;  // [out 0][out 1], 1024-byte aligned tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + NS;
  uint64_t* tmem_full_bar = empty_bar + NS;
  uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::kAccStages;
  uint64_t* res_full_bar = tmem_empty_bar + Cfg::kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full_bar + 2);
  float* stat_red = reinterpret_cast<float*>(smem_epi + Cfg::kEpiBytes + 512);  // STATS: [2][4 warps][8 octets][2], past the barriers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.epi_tma) {
      tma_prefetch_desc(&tmOut);
      if (p.res) tma_prefetch_desc(&tmRes);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's copy is the one used: one arrive per CTA (+ the TMA transaction bytes)
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit
    }
    for (int a = 0; a < Cfg::kAccStages; ++a) {
      mbar_init(&tmem_full_bar[a], 1);                  // multicast tcgen05.commit
      mbar_init(&tmem_empty_bar[a], 2 * kEpiThreads2);  // leader's copy: every epilogue thread of both CTAs
      mbar_init(&res_full_bar[a], 1);                   // residual staging tile landed (TMA transaction bytes)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail (pdl.cuh)
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer (one thread per CTA)
    if (lane == 0) {
      const int cblks = p.Cin / BK2;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        ConvUnit cu = {t, -1};
        if constexpr (TAIL) cu = sched_unit(t, p.tail_full);
        const Tile2 tl = decode_tile(p, cu.tile, n_tiles, pair_tiles, rank);
        const int tw_i = tl.mt % p.tiles_w, th_i = (tl.mt / p.tiles_w) % p.tiles_h, tn_i = tl.mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
        int bcol = tl.n_tile * BN + (int)rank * (BN / 2);
        if constexpr (TAIL) {
          if (cu.half >= 0) bcol = tl.n_tile * BN + cu.half * (BN / 2) + (int)rank * (BN / 4);  // each CTA stages a quarter of the weight tile
        }
        for (int kb = tl.kb0; kb < tl.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const int tap = kb / cblks, cb = kb - tap * cblks;
          int dy = 0, dx = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          bool half_unit = false;
          if constexpr (TAIL) half_unit = cu.half >= 0;
          if (!(p.dbg & 1)) {
            tma2_load_4d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], cb * BK2, w0 + dx, h0 + dy, n0);
            if (half_unit) tma2_load_4d(smem_b + stage * Cfg::kBBytes, &tmB4, &full_bar[stage], kb * BK2, bcol, 0, 0);
            else tma2_load_4d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK2, bcol, p.b_batched ? h0 : 0, p.b_batched ? n0 : 0);
          }
          const uint32_t tx_bytes = half_unit ? 2u * (Cfg::kABytes + Cfg::kBBytes / 2) : 2u * Cfg::kStageBytes;
          if (leader) mbar_expect_tx(&full_bar[stage], (p.dbg & 1) ? 0u : tx_bytes);
          else mbar_arrive_remote(&full_bar[stage], 0);
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA
    if (leader && lane == 0) {
      constexpr uint32_t idesc_full = make_idesc_f16_mn(256, BN);
      constexpr uint32_t idesc_half = make_idesc_f16_mn(256, BN >= 32 ? BN / 2 : 16);
      int stage = 0, iter = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters, ++iter) {
        ConvUnit cu = {t, -1};
        if constexpr (TAIL) cu = sched_unit(t, p.tail_full);
        uint32_t idesc = idesc_full;
        if constexpr (TAIL) idesc = cu.half >= 0 ? idesc_half : idesc_full;
        const Tile2 tl = decode_tile(p, cu.tile, n_tiles, pair_tiles, rank);
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogues of both CTAs have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = tl.kb0; kb < tl.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
          if (!(p.dbg & 2))
#pragma unroll
          for (int k = 0; k < BK2 / 16; ++k)
            tc2_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > tl.kb0 || k > 0) ? 1u : 0u);
          tc2_commit_mc(&empty_bar[stage], 0x3);  // frees the stage in both CTAs
          if (kb == tl.kb1 - 1) tc2_commit_mc(&tmem_full_bar[acc], 0x3);
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      // all remote arrivals on this CTA's barriers must have landed before it may exit
      if (iter > 0) {
        const int last = iter - 1;
        mbar_wait(&tmem_empty_bar[last & 1], (last >> 1) & 1);
        if (iter > 1) mbar_wait(&tmem_empty_bar[(last - 1) & 1], ((last - 1) >> 1) & 1);
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): TMEM lane quadrant = warp % 4
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    int iter = 0;
    if (p.pf_bytes > 0) l2_prefetch_slice(p.pf_ptr, p.pf_bytes, blockIdx.x, gridDim.x, (int)threadIdx.x - 64, kEpiThreads2);
    if constexpr (BN >= 64) {
      if (p.epi_tma) {
        // Accumulator -> registers -> (+bias, +residual) -> fp16 -> 128B-swizzled staging tile -> one TMA tensor store per
        // [128 pixel x 64 channel] chunk.  Measured on B200 (profiles/r01_conv_microbench_v1.txt): the per-thread 16-byte row stores
        // this replaces (32 partially written lines per warp instruction) bounded the 256x256 3x3 layer at 104 us, against
        // 62 us for the same kernel with its stores disabled; the 1x1 layers ran at 24 -> 138 us.
        constexpr int kChunks = BN / 64;
        const bool issuer = threadIdx.x == 64;  // warp 2, lane 0: issues every TMA store / residual load of this CTA
        const bool has_res = p.res != nullptr;
        uint8_t* st_out = smem_epi;  // residual chunks are TMA-loaded INTO the staging tile and updated in place
        auto tile_origin = [&](int tt, int& w0, int& h0, int& n0, int& ncol0) {  // tt: tile index, or schedule unit with TAIL
          ConvUnit cu = {tt, -1};
          if constexpr (TAIL) cu = sched_unit(tt, p.tail_full);
          const Tile2 tl = decode_tile(p, cu.tile, n_tiles, pair_tiles, rank);
          const int tw_i = tl.mt % p.tiles_w, th_i = (tl.mt / p.tiles_w) % p.tiles_h, tn_i = tl.mt / (p.tiles_w * p.tiles_h);
          w0 = tw_i * p.TW;
          h0 = th_i * p.TH;
          n0 = tn_i * p.TN;
          ncol0 = tl.n_tile * BN + (cu.half > 0 ? BN / 2 : 0);
        };
        auto unit_chunks = [&](int tt) {  // 64-channel chunks of unit tt
          if constexpr (TAIL) return sched_unit_chunks(sched_unit(tt, p.tail_full).half, BN);
          else return kChunks;
        };
        auto issue_res = [&](int tt, int c, int buf) {  // issuer only
          int w0, h0, n0, ncol0;
          tile_origin(tt, w0, h0, n0, ncol0);
          mbar_expect_tx(&res_full_bar[buf], kEpiChunkBytes);
          tma_load_4d(st_out + buf * kEpiChunkBytes, &tmRes, &res_full_bar[buf], ncol0 + c * 64, w0, h0, n0);
        };
        if (issuer && has_res) {  // the first two residual chunks of this CTA's tile sequence
          int tt = cluster_id, c = 0;
          for (int k = 0; k < 2 && tt < total_tiles; ++k) {
            issue_res(tt, c, k);
            if constexpr (TAIL) {
              sched_next_chunk(tt, c, p.tail_full, BN, n_clusters);
            } else {
              if (++c == kChunks) {
                c = 0;
                tt += n_clusters;
              }
            }
          }
        }
        uint32_t g = 0;  // running chunk counter of this CTA: staging buffer = g & 1
        const uint32_t sw = (uint32_t)(r & 7);
        for (int t = cluster_id; t < total_tiles; t += n_clusters, ++iter) {
          int w0, h0, n0, ncol0;
          tile_origin(t, w0, h0, n0, ncol0);
          const int acc = iter & 1;
          mbar_wait(&tmem_full_bar[acc], (iter >> 1) & 1);
          tc_fence_after();
          const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
          const int n_chunks = unit_chunks(t);
#pragma unroll 1
          for (int c = 0; c < n_chunks; ++c, ++g) {
            const int buf = (int)(g & 1u);
            const int col = ncol0 + c * 64;
            uint32_t v[64];
            __syncwarp();  // reconverge (issuer-only branch below): the TMEM load is warp-collective (.sync.aligned)
            tc_ld_32x32(taddr_row + c * 64, v);
            tc_ld_32x32(taddr_row + c * 64 + 32, v + 32);
            if (has_res) mbar_wait(&res_full_bar[buf], (g >> 1) & 1u);
            tc_ld_wait();
            if (c == n_chunks - 1) {  // accumulator fully read: hand it back to the MMA thread before the stores
              tc_fence_before();
              mbar_arrive_remote(&tmem_empty_bar[acc], 0);
            }
            // staging tile `buf` is free once the store issued two chunks ago has read it; with a residual that was already
            // awaited before the residual load into this tile was issued (below), and res_full_bar orders its arrival
            if (!has_res) {
              if (issuer) bulk_wait_group_read<1>();
              named_bar_sync(1, kEpiThreads2);
            }
            uint8_t* orow = st_out + buf * kEpiChunkBytes + r * 128;
            const uint8_t* rrow = orow;
            float st_s[8], st_q[8];  // flags 2: per-octet sum / sum of squares of this row's fp16 outputs
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // 8 channels = one 16-byte unit of the 128-byte row, unit index XOR (row & 7)
              float a[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] = __uint_as_float(v[j * 8 + e]);
              if (p.bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col + j * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + j * 8 + 4));
                a[0] += b0.x; a[1] += b0.y; a[2] += b0.z; a[3] += b0.w;
                a[4] += b1.x; a[5] += b1.y; a[6] += b1.z; a[7] += b1.w;
              }
              const uint32_t unit = ((uint32_t)j ^ sw) * 16;
              if (has_res) {
                float rr[8];
                unpack8(*reinterpret_cast<const half8*>(rrow + unit), rr);
                if (p.res_mode == 1) {  // CONV flags 4: out = acc * QuickGELU'(u), u = the tile TMA-loaded through the residual path
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float sg = rcp_ftz(1.f + ex2_ftz(-1.702f * 1.4426950408889634f * rr[e]));  // sigmoid(1.702 u), MUFU ex2 + rcp
                    a[e] *= sg * fmaf(1.702f * rr[e], 1.f - sg, 1.f);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e) a[e] += rr[e];
                }
              }
              const half8 hv = pack8(a);
              *reinterpret_cast<half8*>(orow + unit) = hv;
              if constexpr (STATS) {  // statistics of the ROUNDED values: what the GroupNorm after this conv reads
                float f[8];
                unpack8(hv, f);
                float s0 = 0.f, q0 = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  s0 += f[e];
                  q0 = fmaf(f[e], f[e], q0);
                }
                st_s[j] = s0;
                st_q[j] = q0;
              }
            }
            if constexpr (STATS) {  // 128 rows -> one value per octet: warp shuffles, then the four warps through shared memory
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                st_s[j] = warp_sum(st_s[j]);
                st_q[j] = warp_sum(st_q[j]);
              }
              if (lane == 0) {
                float* dst = stat_red + ((buf * 4 + quad) * 8) * 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  dst[2 * j] = st_s[j];
                  dst[2 * j + 1] = st_q[j];
                }
              }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, kEpiThreads2);
            if (STATS && r < 16) {  // thread (octet j = r / 2, kind = r % 2): fixed-order sum over the four warps
              const float* src = stat_red + (buf * 4 * 8) * 2 + r;
              const float v = (src[0] + src[16]) + (src[32] + src[48]);
              ConvUnit cu2 = {t, -1};
              if constexpr (TAIL) cu2 = sched_unit(t, p.tail_full);
              const Tile2 tl2 = decode_tile(p, cu2.tile, n_tiles, pair_tiles, rank);
              p.epi_stats[((size_t)tl2.mt * (p.Npad / 8) + (size_t)(col / 8 + (r >> 1))) * 2 + (r & 1)] = v;
            }
            if (issuer) {
              if (!(p.dbg & 4)) tma_store_4d(st_out + buf * kEpiChunkBytes, &tmOut, col, w0, h0, n0);
              bulk_commit_group();
              if (has_res) {  // once this store has read tile `buf`, the residual of chunk g + 2 may land in it
                bulk_wait_group_read<0>();
                int tt = t, c2 = c;
                if constexpr (TAIL) {
                  sched_next_chunk(tt, c2, p.tail_full, BN, n_clusters);
                  if (tt < total_tiles) sched_next_chunk(tt, c2, p.tail_full, BN, n_clusters);
                } else {
                  c2 += 2;
                  while (c2 >= kChunks) {
                    c2 -= kChunks;
                    tt += n_clusters;
                  }
                }
                if (tt < total_tiles) issue_res(tt, c2, buf);
              }
            }
          }
        }
        if (issuer) bulk_wait_group<0>();
      }
    }
    if (!TAIL && !p.epi_tma)  // (the TAIL instantiation is only launched with the TMA-store epilogue)
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++iter) {
      const Tile2 tl = decode_tile(p, t, n_tiles, pair_tiles, rank);
      const int tw_i = tl.mt % p.tiles_w, th_i = (tl.mt / p.tiles_w) % p.tiles_h, tn_i = tl.mt / (p.tiles_w * p.tiles_h);
      const int w = tw_i * p.TW + r % p.TW, h = th_i * p.TH + (r / p.TW) % p.TH, n = tn_i * p.TN + r / (p.TW * p.TH);
      const bool row_ok = (n < p.NB) && (h < p.H) && (w < p.W) && !(p.dbg & 4);
      const int ncol0 = tl.n_tile * BN;
      const int acc = iter & 1;
      mbar_wait(&tmem_full_bar[acc], (iter >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      if (p.splits > 1) {
        float* ws = p.ws + ((size_t)tl.split * p.ws_rows + (size_t)tl.mt * BM2 + r) * p.Npad + ncol0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          __syncwarp();
          tc_ld_32x16(taddr_row + c, v);
          tc_ld_wait();
          if (row_ok)
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(ws + c + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        if (p.fuse_reduce) splitk_fused_reduce<BN>(p, tl.mt, tl.n_tile, n_tiles, tl.split, r, n, h, w, row_ok, threadIdx.x == 64);
      } else {
        const int64_t o_off = (int64_t)n * p.out_sn + (int64_t)h * p.out_sh + (int64_t)w * p.out_sw;
        const int64_t r_off = (int64_t)n * p.res_sn + (int64_t)h * p.res_sh + (int64_t)w * p.res_sw;
        // the residual is read ahead of use (one 16-column chunk = 32 B per row in flight while the previous one is
        // converted and stored): un-prefetched, its ~800-cycle global loads serialised 16x per tile and made residual layers
        // epilogue-bound (141 us vs 103 us on the 256x256 layer)
        half8 rnext[2];
        const bool vec_ok = !p.out_f32 && p.out_sc == 1;
        const bool use_res = p.res != nullptr && row_ok && vec_ok;
        if (use_res && ncol0 + 16 <= p.Cout) {
          rnext[0] = ld8(p.res + r_off + ncol0);
          rnext[1] = ld8(p.res + r_off + ncol0 + 8);
        }
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          __syncwarp();
          tc_ld_32x16(taddr_row + c, v);
          half8 rcur[2] = {rnext[0], rnext[1]};
          if (use_res && c + 16 < BN && ncol0 + c + 32 <= p.Cout) {
            rnext[0] = ld8(p.res + r_off + ncol0 + c + 16);
            rnext[1] = ld8(p.res + r_off + ncol0 + c + 24);
          }
          tc_ld_wait();
          const int col = ncol0 + c;
          if (row_ok && col < p.Cout) {
            float a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = __uint_as_float(v[j]);
            if (col + 16 <= p.Cout && vec_ok) {
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const float4 b = *reinterpret_cast<const float4*>(p.bias + col + j);
                  a[j] += b.x; a[j + 1] += b.y; a[j + 2] += b.z; a[j + 3] += b.w;
                }
              }
              if (p.res) {
                float rr[16];
                unpack8(rcur[0], rr);
                unpack8(rcur[1], rr + 8);
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] += rr[j];
              }
              __half* o = reinterpret_cast<__half*>(p.out) + o_off + col;
              st8(o, pack8(a));
              st8(o + 8, pack8(a + 8));
            } else {
              for (int j = 0; j < 16 && col + j < p.Cout; ++j) {
                float x = a[j];
                if (p.bias) x += p.bias[col + j];
                if (p.res) x += __half2float(p.res[r_off + col + j]);
                if (p.out_f32) reinterpret_cast<float*>(p.out)[o_off + (col + j) * p.out_sc] = x;
                else reinterpret_cast<__half*>(p.out)[o_off + (col + j) * p.out_sc] = __float2half_rn(x);
              }
            }
          }
        }
      }
      // this thread's TMEM reads of the accumulator are complete: release it to the leader's MMA thread
      tc_fence_before();
      mbar_arrive_remote(&tmem_empty_bar[acc], 0);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

template <int BN, bool STATS, bool TAIL, bool CO = false>
static int launch_tc2(const ConvTcLaunch& L, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN>;
  constexpr int kSmemMax = Cfg::kSmemBytes + (STATS ? Cfg::kStatBytes : 0);
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    CGD_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BN, STATS, TAIL, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    attr_set.mark();
  }
  const int pair_tiles = (L.m_tiles + 1) / 2;
  const int total = TAIL ? L.tail_units : pair_tiles * L.n_tiles * L.p.splits;  // schedule units
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("CGD_CONV_DBG");
    dbg = e ? atoi(e) : 0;
  }
  ConvTcParams prm = L.p;
  prm.dbg = dbg;
  prm.nstages = (L.p.nstages >= 2 && L.p.nstages <= Cfg::kStages) ? L.p.nstages : Cfg::kStages;
  const int smem = kSmemMax - (Cfg::kStages - prm.nstages) * Cfg::kStageBytes;
  const int clusters = std::min(total, TAIL ? 74 : device_sm_count() / 2);  // one CTA pair per TPC (74 on B200; the TAIL schedule is built for 74)
  CGD_CUDA(launch_pdl(conv_tc2_kernel<BN, STATS, TAIL, CO>, dim3(2 * clusters), dim3(kThreads2), smem, st, L.tmA, L.tmB2, L.tmOut, L.tmRes, L.tmB4, prm,
                      L.n_tiles, pair_tiles, total));
  return 0;
}

template <bool STATS, bool TAIL>
static int launch_tc2_wide(const ConvTcLaunch& L, cudaStream_t st) {  // the instantiations that need the TMA-store epilogue (BN >= 64)
  switch (L.BN) {
    case 64: if constexpr (!TAIL) return launch_tc2<64, STATS, false>(L, st); else break;
    case 128: return launch_tc2<128, STATS, TAIL>(L, st);
    case 192: if constexpr (!TAIL) return launch_tc2<192, STATS, false>(L, st); else break;
    case 256: return launch_tc2<256, STATS, TAIL>(L, st);
    default: break;
  }
  set_error("conv: BN %d has no %s%s instantiation", L.BN, STATS ? "epilogue-statistics " : "", TAIL ? "split-tail " : "");
  return -1;
}

int conv_tc2_launch(const ConvTcLaunch& L, cudaStream_t st) {
  const bool stats = L.p.epi_stats != nullptr, tail = L.tail_units > 0;  // prepare() has checked the TMA-store epilogue for both
  if (stats && tail) return launch_tc2_wide<true, true>(L, st);
  if (stats) return launch_tc2_wide<true, false>(L, st);
  if (tail) return launch_tc2_wide<false, true>(L, st);
  if (L.co_resident) {
    switch (L.BN) {
      case 16: return launch_tc2<16, false, false, true>(L, st);
      case 32: return launch_tc2<32, false, false, true>(L, st);
      case 64: return launch_tc2<64, false, false, true>(L, st);
      case 128: return launch_tc2<128, false, false, true>(L, st);
      case 192: return launch_tc2<192, false, false, true>(L, st);
      case 256: return launch_tc2<256, false, false, true>(L, st);
      default: set_error("conv: unsupported BN %d", L.BN); return -1;
    }
  }
  switch (L.BN) {
    case 16: return launch_tc2<16, false, false>(L, st);
    case 32: return launch_tc2<32, false, false>(L, st);
    case 64: return launch_tc2<64, false, false>(L, st);
    case 128: return launch_tc2<128, false, false>(L, st);
    case 192: return launch_tc2<192, false, false>(L, st);
    case 256: return launch_tc2<256, false, false>(L, st);
    default: set_error("conv: unsupported BN %d", L.BN); return -1;
  }
}

}  // namespace cgd
