#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fill.py tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_goldens.py tests/test_gpu_api.py tests/test_gpu_flatten.py tests/test_gpu_boundary.py -x -q 2>&1 | tail -2
for s in 900 2048 4096 8192; do PIXIE_CUDA_LIB=build/pixie_cuda_base.so TIGER_CLEAR=1 timeout 120 python tools/time_tiger.py $s; TIGER_CLEAR=1 timeout 120 python tools/time_tiger.py $s; done
for i in 1 2; do PIXIE_CUDA_LIB=build/pixie_cuda_base.so timeout 120 python tools/time_icons.py | tail -1; timeout 120 python tools/time_icons.py | tail -1; done
PIXIE_CUDA_LIB=build/pixie_cuda_base.so python tools/time_e2e_host.py 2>&1 | tail -1
python tools/time_e2e_host.py 2>&1 | tail -1
