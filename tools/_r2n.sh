python -m pytest tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_fill.py tests/test_gpu_goldens.py tests/test_gpu_flatten.py tests/test_gpu_boundary.py -x -q 2>&1 | tail -4
python tools/time_tiger.py
python tools/time_tiger.py 2048
python tools/time_tiger.py 8192
