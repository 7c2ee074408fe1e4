python -m pytest tests/test_gpu_tiger.py -x -q 2>&1 | tail -5
python -m pytest tests/test_gpu_flatten.py -x -q 2>&1 | tail -30
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/time_tiger.py
