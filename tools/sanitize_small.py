"""Small workloads for compute-sanitizer (memcheck / racecheck): tiger at 600^2 (cut rows need h >= 256, >= 16 fills), a few
icons on layers, run_cleared on a dirty canvas, the float blend modes, gradients."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import tiger_arrays  # noqa: E402
from pixie_b200 import device as dev, synth  # noqa: E402
from pixie_b200.device import FillBatch  # noqa: E402

dev.init(0)
os.environ.setdefault("PIXIE_CUDA_CUT", "30")  # cut many rows
for size in (600, 2100):
    arrays = tiger_arrays(size)
    img = dev.DeviceImage(size, size).upload(synth.random_premultiplied(size, size, 3))
    cl = dev.CmdList(size, size, 1, arrays)
    cl.run(img, clear=True)
    print("tiger", size, "checksum", img.checksum())
b = FillBatch()
for i in range(6):
    synth.icon_fills(i, 200, i, b)
img = dev.DeviceImage(200, 200, 6)
cl = dev.CmdList(200, 200, 6, b.arrays())
cl.run(img, clear=True)
print("icons checksum", img.checksum())
n = 256
dst = dev.DeviceImage(n, n).upload(synth.random_premultiplied(n, n, 1))
src = dev.DeviceImage(n, n).upload(synth.random_premultiplied(n, n, 2))
mask = dev.DeviceImage(n, n, a8=True).upload(synth.coverage_mask(n, n, 3))
for mode in (3, 6, 8, 12, 13, 14, 15):
    dev.blend_rect_masked(dst, src, mask, 0, 0, mode)
stops = [(0.0, (1.0, 0.2, 0.1, 1.0)), (0.35, (0.1, 0.9, 0.3, 0.4)), (1.0, (0.2, 0.1, 1.0, 0.85))]
dev.fill_gradient(dst, 3, [(10.0, 20.0), (200.0, 180.0)], stops, 1.0)
dev.sync()
print("done", dst.checksum())
