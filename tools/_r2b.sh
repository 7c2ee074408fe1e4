bash tools/gpu_round.sh r2a
PIXIE_CUDA_TC_DEBUG=1 python tools/time_blur.py 16384 32 > gpurun_out/r2a_blur_dbg.log 2>&1
python tools/time_blur.py 16384 29,30,31,32 > gpurun_out/r2a_blur_radii.log 2>&1
for b in 1 2 4 8; do PIXIE_CUDA_BANDS=$b python tools/time_tiger.py; done > gpurun_out/r2a_tiger_bands.log 2>&1
cat gpurun_out/r2a_blur_dbg.log gpurun_out/r2a_blur_radii.log gpurun_out/r2a_tiger_bands.log
