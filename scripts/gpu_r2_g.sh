#!/bin/bash
# Round 2, call G: stand-alone probe of what bounds a GroupNorm-apply-like stream (scripts/probes/gn_stream_probe.cu)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/gn_probe scripts/probes/gn_stream_probe.cu && /tmp/gn_probe | tee gpurun_out/r02_gn_stream_probe_v1.txt
