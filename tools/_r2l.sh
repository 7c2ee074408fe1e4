export PIXIE_CUDA_TC=16
timeout 300 python -m pytest tests/test_gpu_blend_blur.py -x -q -k "blur" 2>&1 | tail -5
timeout 120 python tools/time_blur.py 16384 32 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_baseline_sizes.py -x -q 2>&1 | tail -3
