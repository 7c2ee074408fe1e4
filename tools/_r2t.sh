python -m pytest tests/test_gpu_flatten.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
python tools/time_flatten.py 4096 | tail -3
