#!/bin/bash
for so in pixie_b200/pixie_cuda.so build/pixie_cuda_a3.so build/pixie_cuda_a4.so build/pixie_cuda_a6.so; do echo $so; PIXIE_CUDA_LIB=$so timeout 120 python tools/time_blend.py 3,6 2>&1 | tail -2; done
PIXIE_CUDA_LIB=build/pixie_cuda_a4.so timeout 600 python -m pytest tests/test_gpu_blend_blur.py -x -q -k blend 2>&1 | tail -2
