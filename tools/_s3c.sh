#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_blend_blur.py -x -q -k "blend" 2>&1 | tail -3
PIXIE_CUDA_LIB=build/pixie_cuda_base.so timeout 120 python tools/time_blend.py 3,8,12,13,14,15 2>&1 | tail -6
timeout 120 python tools/time_blend.py 3,8,12,13,14,15 2>&1 | tail -6
