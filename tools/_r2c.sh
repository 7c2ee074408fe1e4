python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
python bench.py --no-extras > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err
python tools/time_tiger.py
PIXIE_CUDA_BANDS=3 python tools/time_tiger.py
PIXIE_CUDA_BANDS=6 python tools/time_tiger.py
cat gpurun_out/r2c_bench.json | head -c 2500
