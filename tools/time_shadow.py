"""shadow 16384^2 (offset 8,8, spread 4, blur 32) and its pieces."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, host  # noqa: E402


dev.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(1)
src = dev.DeviceImage(n, n)
dst = dev.DeviceImage(n, n)
tile = rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8)
src.upload(np.tile(tile, (n // 1024, n // 1024, 1)))


def t(fn, reps=10):
    """device time between the library's own events (on its stream)"""
    for _ in range(3):
        fn()
    dev.sync()
    dev.timer_begin()
    for _ in range(reps):
        fn()
    return dev.timer_end() / reps


import time
for radius, spread in [(0, 4), (32, 4), (32, 0), (8, 4), (64, 1), (0, 4)]:
    lut = host.gaussianKernel(radius)
    print("shadow %d^2 r=%d spread=%d   %.3f ms" % (n, radius, spread, t(lambda: dev.shadow(src, dst, 8, 8, spread, lut, radius, 0xC8000000))))
