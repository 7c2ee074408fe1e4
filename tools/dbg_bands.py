"""Debug: repeated runs of a banded resident list must give the same canvas."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import tiger_arrays
from pixie_b200 import device as dev
size = 4096
dev.init(0)
if len(sys.argv) > 1 and sys.argv[1] == "prof":
    dev.set_profiling(True)
arrays = tiger_arrays(size)
img = dev.DeviceImage(size, size)
cl = dev.CmdList(size, size, 1, arrays)
for it in range(4):
    img.fill(0)
    c = cl.run(img, count_covered=True)
    print(os.environ.get("PIXIE_CUDA_BANDS", "default"), it, c, img.checksum(), cl.info())
for it in range(3):
    img.fill(0)
    cl.run(img)
    dev.sync()
    print("nocount", it, img.checksum())
