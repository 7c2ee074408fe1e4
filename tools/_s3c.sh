#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_draw.py tests/test_gpu_goldens.py tests/test_gpu_api.py -x -q 2>&1 | tail -2
PIXIE_CUDA_LIB=build/pixie_cuda_base.so timeout 200 python tools/time_paint.py 2>&1 | tail -6
timeout 200 python tools/time_paint.py 2>&1 | tail -6
PIXIE_CUDA_LIB=build/pixie_cuda_base.so timeout 200 python tools/time_draw.py 2>&1 | grep -i grad
timeout 200 python tools/time_draw.py 2>&1 | grep -i grad
