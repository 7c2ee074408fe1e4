"""Determinism under repetition: the tiger at several sizes, 300 runs each, checksum after every run (the plan stage's
per-sample-line split hands spans between warps through global memory; the raster kernel's tickets are dynamic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import tiger_arrays  # noqa: E402
from pixie_b200 import device as dev  # noqa: E402

dev.init(0)
for size in (900, 2048, 4096):
    arrays = tiger_arrays(size)
    img = dev.DeviceImage(size, size)
    cl = dev.CmdList(size, size, 1, arrays)
    cl.run(img, clear=True)
    want = img.checksum()
    bad = 0
    for it in range(300):
        cl.run(img, clear=True)
        if img.checksum() != want:
            bad += 1
    print(f"tiger {size}^2: 300 runs, {bad} checksums differ from the first ({want})")
    assert bad == 0
print("ok")
