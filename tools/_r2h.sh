python -m pytest tests/test_gpu_draw.py tests/test_gpu_goldens.py tests/test_gpu_api.py tests/test_gpu_blend_blur.py -x -q 2>&1 | tail -5
python tools/time_paint.py
python tools/time_blend.py 3,6
