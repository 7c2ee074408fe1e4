#!/bin/bash
# One gpurun call: parity tests, bench (both arms), ncu launch list, ncu --set full of the hot kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_cur_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'plan_|raster_kernel|partition_kernel' --launch-skip 10 -c 5 -f -o gpurun_out/prof_tiger_cur python tools/prof_kernels.py tiger > gpurun_out/ncu_tiger.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'blur_mma' --launch-skip 2 -c 2 -f -o gpurun_out/prof_blur_cur python tools/prof_kernels.py blur > gpurun_out/ncu_blur.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'blur_mma_a8|spread_' --launch-skip 4 -c 4 -f -o gpurun_out/prof_shadow_cur python tools/prof_kernels.py shadow > gpurun_out/ncu_shadow.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'blend_rect' --launch-skip 3 -c 1 -f -o gpurun_out/prof_blend_cur python tools/prof_kernels.py blend > gpurun_out/ncu_blend.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'draw_smooth|gradient_kernel|minify' -c 6 -f -o gpurun_out/prof_draw_cur python tools/prof_kernels.py draw > gpurun_out/ncu_draw.log 2>&1
tail -2 gpurun_out/bench_cur.err
head -c 1500 gpurun_out/bench_cur.json
