"""Blur r=32 (and other radii) at 16384^2: the fused tcgen05 kernel vs PIXIE_CUDA_BLUR=mma (set in the environment)."""
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixie_b200 import device as dev, host, synth

dev.init(0)
dev.set_profiling(True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
radii = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [32]
img0 = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 0xB10B), (n // 512, 1, 1)))
img = dev.DeviceImage(n, n)
for r in radii:
    lut = host.gaussianKernel(r)
    ms = []
    for it in range(6):
        img.copy_from(img0)
        dev.timer_begin()
        dev.blur(img, lut, r, 0)
        ms.append(dev.timer_end())
    t = statistics.median(ms[2:])
    print(f"blur r={r} {n}^2 [{os.environ.get('PIXIE_CUDA_BLUR', 'tc')}]: {t:.3f} ms  {n * n * 8 / t / 1e6:.0f} GB/s  frac {n * n * 8 / t / 1e6 / 6550.4:.3f}  all {[round(v, 3) for v in ms]}")
