#!/bin/bash
python -m pytest tests/test_gpu_fill.py tests/test_gpu_tiger.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
PIXIE_CUDA_LIB=build/pixie_cuda_base.so python tools/time_tiger.py
python tools/time_tiger.py
PIXIE_CUDA_LIB=build/pixie_cuda_base.so python tools/time_icons.py 2>&1 | tail -2
python tools/time_icons.py 2>&1 | tail -2
