#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fill.py tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_goldens.py -x -q 2>&1 | tail -2
for so in build/pixie_cuda_base.so pixie_b200/pixie_cuda.so; do
PIXIE_CUDA_LIB=$so python tools/time_tiger.py
PIXIE_CUDA_LIB=$so python tools/time_icons.py 2>&1 | tail -1
done
