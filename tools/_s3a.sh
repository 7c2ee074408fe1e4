#!/bin/bash
# source-level capture of raster_kernel + plan_kernel on the icon batch (per-line instruction counts)
O=gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -f -k regex:'raster_kernel|plan_kernel' --launch-skip 2 -c 2 -o $O/s3i python tools/prof_kernels.py icons > $O/s3i.log 2>&1
ncu -i $O/s3i.ncu-rep --page source --csv --print-source sass > $O/s3i_sass.csv 2>$O/s3i_err.log
rm -f $O/s3i.ncu-rep
tail -3 $O/s3i.log
