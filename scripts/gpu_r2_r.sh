#!/bin/bash
# Round 2, call R (2 GPU-minutes left): the 577-token attention case alone, then the ViT-L/14@336px full step if time remains.
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_attention.py -q -m gpu -p no:cacheprovider -k "t577" 2>&1 | tail -3 | tee gpurun_out/r02_pytest_attn577_v1.log
CGD_TEST_FIRST_RUN=1 timeout 100 python -m pytest tests/test_gpu_zz_first_run.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -4 | tee gpurun_out/r02_pytest_vit336_v1.log
