"""Gradient paint through a shape mask at 8192^2: the fused composite (pixie_cuda_fill_gradient_masked) against the
three separate passes (fillGradient + blend_rect_masked; the mask itself is rendered by fillPath in both cases)."""
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, host, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev.init(0)
dst0 = np.tile(synth.random_premultiplied(512, n, 1), (n // 512, 1, 1))
dst = dev.DeviceImage(n, n).upload(dst0)
mask = dev.DeviceImage(n, n)
p = host.newPath()
p.ellipse(n / 2, n / 2, n * 0.3, n * 0.22)  # covers ~21 % of the canvas
dev.fill_segments(mask, host.fill_segments(p), 0xFFFFFFFF, 0, 0)
fill = dev.DeviceImage(n, n)
stops = [(0.0, (1.0, 0.2, 0.1, 1.0)), (0.35, (0.1, 0.9, 0.3, 0.4)), (1.0, (0.2, 0.1, 1.0, 0.85))]
for kind, handles, name in ((3, [(n * 0.2, n * 0.3), (n * 0.9, n * 0.7)], "linear"),
                            (4, [(n * 0.5, n * 0.5), (n * 0.9, n * 0.5), (n * 0.5, n * 0.95)], "radial"),
                            (5, [(n * 0.5, n * 0.5), (n * 0.9, n * 0.5), (n * 0.5, n * 0.95)], "angular")):
    for mode, mname in ((0, "Normal"), (7, "Overlay")):
        ta, tb = [], []
        for it in range(5):
            dev.timer_begin()
            dev.fill_gradient(fill, kind, handles, stops, 1.0)
            dev.blend_rect_masked(dst, fill, mask, 0, 0, mode)
            ta.append(dev.timer_end())
            dev.timer_begin()
            dev.fill_gradient_masked(dst, mask, kind, handles, stops, 1.0, mode)
            tb.append(dev.timer_end())
        print(f"{name:8s} {mname:8s} {n}^2: separate {statistics.median(ta[1:]):.3f} ms  fused {statistics.median(tb[1:]):.3f} ms")
