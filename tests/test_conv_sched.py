"""CPU: the tile schedule of the pair conv kernel with a split last wave (csrc/conv_sched.cuh), compiled as plain C++ and checked
exhaustively: every output chunk of every tile is produced exactly once, by the cluster the round-robin assigns it to, and the
look-ahead iterator the epilogue uses for its residual prefetch visits the chunks in exactly the order they are consumed."""
import os
import shutil
import subprocess

import pytest

SRC = r'''
#include <cstdio>
#include <vector>
#include <set>
#include "conv_sched.cuh"
using namespace cgd;
int main() {
  long checked = 0;
  for (int BN : {128, 256})
    for (int ncl : {1, 3, 37, 74})
      for (int total = 1; total <= 700; ++total) {
        const bool tail = sched_tail_pays(total, ncl);
        const int F = tail ? sched_full_tiles(total, ncl) : total;
        const int units = tail ? sched_total_units(total, ncl) : total;
        if (tail && (F % ncl != 0 || units - F > ncl || units - F != 2 * (total - F))) { printf("bad split %d %d\n", total, ncl); return 1; }
        std::set<long> seen;  // (tile, 64-channel chunk of the tile)
        int max_units_per_cluster = 0;
        for (int cl = 0; cl < ncl; ++cl) {
          // consumption order of cluster cl
          std::vector<std::pair<int,int>> order;
          int n_units = 0;
          for (int u = cl; u < units; u += ncl, ++n_units) {
            const ConvUnit cu = sched_unit(u, F);
            if (cu.tile < 0 || cu.tile >= total) { printf("tile out of range\n"); return 1; }
            const int chunks = sched_unit_chunks(cu.half, BN);
            for (int c = 0; c < chunks; ++c) {
              const int chunk_in_tile = (cu.half > 0 ? BN / 128 : 0) + c;
              const long key = (long)cu.tile * 16 + chunk_in_tile;
              if (!seen.insert(key).second) { printf("chunk produced twice: total %d ncl %d tile %d\n", total, ncl, cu.tile); return 1; }
              order.push_back({u, c});
            }
          }
          if (n_units > max_units_per_cluster) max_units_per_cluster = n_units;
          // the look-ahead iterator reproduces that order
          int u = cl, c = 0;
          for (size_t k = 0; k < order.size(); ++k) {
            if (u != order[k].first || c != order[k].second) { printf("look-ahead diverges: total %d ncl %d\n", total, ncl); return 1; }
            sched_next_chunk(u, c, F, BN, ncl);
          }
          if (u < units) { printf("look-ahead ends early\n"); return 1; }
        }
        if ((long)seen.size() != (long)total * (BN / 64)) { printf("coverage %ld vs %ld (total %d ncl %d)\n", (long)seen.size(), (long)total * (BN / 64), total, ncl); return 1; }
        if (tail) {  // the point of it: one more unit than the full waves, and that unit is a half
          if (max_units_per_cluster != F / ncl + 1) { printf("units per cluster %d\n", max_units_per_cluster); return 1; }
        }
        ++checked;
      }
  // the dominant layer of cfg2: 256 tiles on 74 pairs -> 222 whole tiles + 68 halves
  if (!sched_tail_pays(256, 74) || sched_full_tiles(256, 74) != 222 || sched_total_units(256, 74) != 290) return 2;
  if (sched_tail_pays(296, 74) || sched_tail_pays(60, 74) || sched_tail_pays(74 * 3 + 40, 74)) return 3;  // exact waves, single wave, tail too long
  printf("ok %ld\n", checked);
  return 0;
}
'''


def test_tail_split_schedule(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sched.cpp"
    src.write_text(SRC)
    exe = tmp_path / "sched"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(root, "clip_guided_diffusion_b200", "csrc"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
