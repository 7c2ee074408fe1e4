#!/bin/bash
for v in tk s96 s200 s100000; do echo $v; PIXIE_CUDA_LIB=build/pixie_cuda_$v.so TIGER_CLEAR=1 timeout 120 python tools/time_tiger.py 2>&1 | tail -3 | grep -v raster.tickets; PIXIE_CUDA_LIB=build/pixie_cuda_$v.so timeout 120 python tools/time_icons.py 2>&1 | tail -3 | grep -v "raster tickets"; done
