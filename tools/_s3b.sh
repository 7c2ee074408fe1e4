#!/bin/bash
python -m pytest tests/test_gpu_fill.py tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_goldens.py -x -q 2>&1 | tail -3
for so in build/pixie_cuda_base.so pixie_b200/pixie_cuda.so; do
PIXIE_CUDA_LIB=$so python tools/time_tiger.py
PIXIE_CUDA_LIB=$so python tools/time_icons.py 2>&1 | tail -1
done
O=gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -f -k regex:'raster_kernel' --launch-skip 2 -c 1 -o $O/s3a python tools/prof_kernels.py tiger > $O/s3a.log 2>&1
ncu -i $O/s3a.ncu-rep --page source --csv --print-source sass > $O/s3a_sass.csv 2>>$O/s3a_err.log
rm -f $O/s3a.ncu-rep
