#!/bin/bash
# source-level capture of raster_kernel + plan kernels on the tiger (per-line instruction counts)
O=gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -f -k regex:'raster_kernel|plan_kernel|plan_light' --launch-skip 6 -c 3 -o $O/s3a python tools/prof_kernels.py tiger > $O/s3a.log 2>&1
ncu -i $O/s3a.ncu-rep --page source --csv --print-source cuda > $O/s3a_cuda.csv 2>$O/s3a_err.log
ncu -i $O/s3a.ncu-rep --page source --csv --print-source sass > $O/s3a_sass.csv 2>>$O/s3a_err.log
rm -f $O/s3a.ncu-rep
ls -la $O/s3a*
head -c 600 $O/s3a_cuda.csv
