python -m pytest tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_fill.py tests/test_gpu_goldens.py -x -q 2>&1 | tail -3
python tools/time_tiger.py
python tools/time_icons.py | tail -1
