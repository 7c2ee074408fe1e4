PIXIE_CUDA_TRACE=1 python tools/time_e2e_host.py 2>&1 | tail -12
