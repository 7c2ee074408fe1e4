#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_blend_blur.py tests/test_gpu_baseline_sizes.py tests/test_gpu_draw.py tests/test_gpu_fill.py -x -q 2>&1 | tail -3
timeout 120 python tools/time_blend.py 2>&1 | tail -8
