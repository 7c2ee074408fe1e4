#!/bin/bash
O=gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -f -k regex:'raster_kernel' --launch-skip 2 -c 1 -o $O/s3a python tools/prof_kernels.py tiger > $O/s3a.log 2>&1
ncu -i $O/s3a.ncu-rep --page source --csv --print-source sass > $O/s3a_sass.csv 2>>$O/s3a_err.log
rm -f $O/s3a.ncu-rep
