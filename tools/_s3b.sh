#!/bin/bash
for c in 50 75 100 130 170 250 1000000; do echo "cut $c"; PIXIE_CUDA_CUT=$c PIXIE_CUDA_LIB=build/pixie_cuda_tk.so python tools/time_tiger.py 2>&1 | tail -2 | head -1; PIXIE_CUDA_CUT=$c TIGER_CLEAR=1 python tools/time_tiger.py; done
