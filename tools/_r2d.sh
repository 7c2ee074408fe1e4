for b in 1 2 4; do PIXIE_CUDA_BANDS=$b python tools/dbg_bands.py; done 2>&1 | tail -30
python tools/dbg_bands.py prof 2>&1 | tail -8
python -m pytest tests/test_gpu_flatten.py -x -q 2>&1 | tail -30
