#!/bin/bash
for c in 50 70 85 100 120 150; do echo "cut $c"; export PIXIE_CUDA_CUT=$c; PIXIE_CUDA_LIB=build/pixie_cuda_tk.so python tools/time_tiger.py 2>&1 | tail -2 | head -1; TIGER_CLEAR=1 python tools/time_tiger.py;  TIGER_CLEAR=1 python tools/time_tiger.py 2048; TIGER_CLEAR=1 python tools/time_tiger.py 8192; done
