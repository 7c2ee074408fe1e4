python -m pytest tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_fill.py -x -q 2>&1 | tail -3
python tools/time_tiger.py
for v in l12v64 l12v32 l8v32 l16v32; do PIXIE_CUDA_LIB=build/ab/$v.so python tools/time_tiger.py; done
python tools/time_icons.py | tail -1
for v in l12v64 l12v32 l8v32 l16v32; do PIXIE_CUDA_LIB=build/ab/$v.so python tools/time_icons.py | tail -1; done
