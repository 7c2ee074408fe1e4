#!/bin/bash
for cap in 64 80 96; do for so in pixie_b200/pixie_cuda.so build/pixie_cuda_s96.so build/pixie_cuda_s80.so; do echo "cap $cap $so"; PIXIE_CUDA_SMEM_CAP=$cap PIXIE_CUDA_LIB=$so TIGER_CLEAR=1 timeout 120 python tools/time_tiger.py; PIXIE_CUDA_SMEM_CAP=$cap PIXIE_CUDA_LIB=$so timeout 120 python tools/time_icons.py | tail -1; done; done
