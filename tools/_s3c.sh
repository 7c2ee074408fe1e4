#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_blend_blur.py tests/test_gpu_baseline_sizes.py -x -q -k "blur or shadow" 2>&1 | tail -2
for i in 1 2; do PIXIE_CUDA_LIB=build/pixie_cuda_base.so timeout 120 python tools/time_blur.py 2>&1 | tail -1; timeout 120 python tools/time_blur.py 2>&1 | tail -1; done
