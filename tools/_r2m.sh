python -m pytest tests/test_gpu_flatten.py tests/test_gpu_api.py tests/test_gpu_tiger.py -x -q 2>&1 | tail -5
python - <<'PY'
import sys; sys.path.insert(0,'.')
from pixie_b200 import device as dev
import bench, json
dev.init(0)
print(json.dumps(bench.extras_flatten(dev))[:1800])
PY
