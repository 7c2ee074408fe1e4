#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fill.py tests/test_gpu_tiger.py tests/test_gpu_fuzz.py tests/test_gpu_goldens.py tests/test_gpu_api.py tests/test_gpu_boundary.py tests/test_gpu_flatten.py -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','roofline','e2e','gpu_launches')})"
python tools/time_e2e_host.py 2>&1 | tail -2
