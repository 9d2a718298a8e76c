// conv_sched.cuh -- tile schedule of the persistent pair kernel with a split last wave (conv_tc2.cu, TAIL instantiation).
//
// 256 output tiles on 74 CTA pairs are 3.46 waves: the last wave keeps 34 pairs busy and 40 idle (86.5 % of the tile slots of four
// waves).  With the tail split, the tiles of the last partial wave are cut in two along the channel dimension (two BN/2-wide halves):
// 222 full tiles + 68 half tiles -> every pair gets three full tiles and at most one half tile, 3.5 tile-times instead of 4.
// Pure index arithmetic, shared by the device code and a host unit test (tests/test_conv_sched.py compiles this header with g++).
#pragma once

#if defined(__CUDACC__)
#define CGD_HD __host__ __device__ __forceinline__
#else
#define CGD_HD inline
#endif

namespace cgd {

struct ConvUnit {
  int tile;  // index into the (channel tile fastest, pixel-pair tile) enumeration of full tiles
  int half;  // -1: the whole BN-wide tile; 0 / 1: its lower / upper BN/2 channels
};

// number of tiles that stay whole: all complete waves.  The rest (fewer than n_clusters tiles) is split.
CGD_HD int sched_full_tiles(int total_tiles, int n_clusters) { return (total_tiles / n_clusters) * n_clusters; }
CGD_HD int sched_total_units(int total_tiles, int n_clusters) {
  const int F = sched_full_tiles(total_tiles, n_clusters);
  return F + 2 * (total_tiles - F);
}
// the split pays when the halves of the tail still fit one wave
CGD_HD bool sched_tail_pays(int total_tiles, int n_clusters) {
  const int T = total_tiles - sched_full_tiles(total_tiles, n_clusters);
  return total_tiles > n_clusters && T > 0 && 2 * T <= n_clusters;
}
CGD_HD ConvUnit sched_unit(int u, int F) {
  ConvUnit r;
  if (u < F) {
    r.tile = u;
    r.half = -1;
  } else {
    const int h = u - F;
    r.tile = F + (h >> 1);
    r.half = h & 1;
  }
  return r;
}
CGD_HD int sched_unit_width(int half, int BN) { return half < 0 ? BN : BN / 2; }
CGD_HD int sched_unit_chunks(int half, int BN) { return sched_unit_width(half, BN) / 64; }
// next 64-channel output chunk in the order ONE cluster produces them: chunks of a unit, then the cluster's next unit
CGD_HD void sched_next_chunk(int& u, int& c, int F, int BN, int n_clusters) {
  if (++c >= sched_unit_chunks(sched_unit(u, F).half, BN)) {
    c = 0;
    u += n_clusters;
  }
}

}  // namespace cgd
