import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev
import bench
dev.init(0)
t0 = time.time()
print(bench.icons_batch(dev, 0, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 1024), "host gen+run %.1fs" % (time.time() - t0))
