#!/bin/bash
# One gpurun call: parity tests, bench (both arms), ncu launch lists, ncu --set full of the hot kernels.
# Usage: tools/gpu_round.sh [tag] [parts]   (outputs under gpurun_out/<tag>_*; parts = any of: test bench lists ncu)
# The .ncu-rep files are summarised on the box (tools/summarize_ncu.py) and deleted: gpurun_out/ merges back only
# below 64 MiB.
set -x
T=${1:-cur}
PARTS=${2:-"test bench lists ncu"}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
FULL="$NCU --set full --import-source on -f"
TENSOR="--metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum,sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active"
cap() {  # cap <name> <kernel regex> <skip> <count> <prof_kernels target> [extra ncu args]
  local name=$1 rx=$2 skip=$3 cnt=$4 target=$5; shift 5
  timeout 600 $FULL "$@" -k regex:"$rx" --launch-skip $skip -c $cnt -o $O/${T}_prof_$name python tools/prof_kernels.py $target > $O/ncu_$name.log 2>&1
  python tools/summarize_ncu.py $O/${T}_prof_$name.ncu-rep $O/${T}_$name > /dev/null 2>&1
  rm -f $O/${T}_prof_$name.ncu-rep
}
if [[ $PARTS == *test* ]]; then
  python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
  tail -3 $O/${T}_pytest_gpu.log
fi
if [[ $PARTS == *bench* ]]; then
  python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
  python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_ref.json 2>/dev/null
  tail -2 $O/${T}_bench.err
  head -c 1500 $O/${T}_bench.json
fi
if [[ $PARTS == *lists* ]]; then
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${T}_launches_tiger.csv python bench.py --steps 3 --warmup 3 --no-extras > $O/ncu_bench.log 2>&1
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${T}_launches_icons.csv python tools/time_icons.py > $O/ncu_icons_l.log 2>&1
  $NCU --metrics gpu__time_duration.sum -c 100 --csv --log-file $O/${T}_launches_blur.csv python tools/prof_kernels.py blur > $O/ncu_blur_l.log 2>&1
  $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $O/${T}_launches_flatten.csv python tools/prof_kernels.py flatten > $O/ncu_flatten_l.log 2>&1
fi
if [[ $PARTS == *ncu* ]]; then
  cap tiger 'plan_|raster_kernel|partition_kernel|count_kernel|band_' 12 9 tiger
  cap blur_tc 'blur_tc' 1 1 blur $TENSOR
  cap shadow 'blur_mma_a8|spread_' 4 4 shadow
  cap blend 'blend_rect' 0 15 blend
  cap icons 'plan_|raster_kernel|partition_kernel' 5 5 icons
  cap draw 'draw_smooth|gradient_kernel|minify' 0 6 draw
  cap flatten 'resolve_kernel|count_kernel_f|shape_first|emit_kernel_f|stroke_count|stroke_emit|bounds_kernel|scan_' 22 22 flatten
  cap paint 'gradient_blend_kernel' 1 3 paint
fi
du -sh $O
