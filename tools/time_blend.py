"""Masked blend at 8192^2 (BASELINE config 3) per mode: ms and fraction of the measured HBM peak (13 B/px)."""
import statistics
import sys
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, synth, common
n = 8192
dev.init(0)
dev.set_profiling(True)
dst0 = np.tile(synth.random_premultiplied(512, n, 1), (n // 512, 1, 1))
dst = dev.DeviceImage(n, n).upload(dst0)
src = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 2), (n // 512, 1, 1)))
mask = dev.DeviceImage(n, n, a8=True).upload(np.tile(synth.coverage_mask(512, n, 3), (n // 512, 1)))
modes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 3, 6, 8, 12, 13, 14, 15]
for mode in modes:
    ts = []
    for it in range(5):
        dev.blend_rect_masked(dst, src, mask, 0, 0, mode)
        ts.append(dev.profile_read(dev.PROF_BLEND))
    t = statistics.median(ts[1:])
    print(f"{common.BLEND_MODE_NAMES[mode]:18s} {t:.4f} ms  frac {13 * n * n / t / 1e6 / 6550.4:.3f}")
