#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fill.py tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_goldens.py -x -q 2>&1 | tail -2
PIXIE_CUDA_LIB=build/pixie_cuda_tk.so python tools/time_tiger.py 2>&1 | tail -2
for i in 1 2; do
PIXIE_CUDA_LIB=build/pixie_cuda_base.so TIGER_CLEAR=1 python tools/time_tiger.py
TIGER_CLEAR=1 python tools/time_tiger.py
done
PIXIE_CUDA_LIB=build/pixie_cuda_base.so python tools/time_icons.py | tail -1
python tools/time_icons.py | tail -1
