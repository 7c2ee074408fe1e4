"""Device timings of the draw / minify / gradient kernels at BASELINE config-3 scale (CUDA events)."""
import math
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, host, synth  # noqa: E402

f = np.float32
dev.init(0)
n = int(os.environ.get("DRAW_N", "8192"))
tile = synth.random_premultiplied(512, n, 1)
dst0 = dev.DeviceImage(n, n).upload(np.tile(tile, (n // 512, 1, 1)))
src = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 2), (n // 512, 1, 1)))
dst = dev.DeviceImage(n, n)


def timed(fn, reps=4):
    ts = []
    for it in range(reps):
        dst.copy_from(dst0)
        dev.timer_begin()
        fn()
        t = dev.timer_end()
        if it:
            ts.append(t)
    return statistics.median(ts)


def report(name, ms, nbytes):
    print(f"{name:44s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s")


T = host.translate
rot = host.matmul(T(f(n / 2), f(-n / 5)), host.rotate(f(0.5)))
for name, m, mode in [("draw frac translate Normal", T(f(10.5), f(3.25)), 0), ("draw rotate 0.5 rad Normal", rot, 0),
                      ("draw rotate Overwrite", rot, 17), ("draw rotate Mask", rot, 16), ("draw rotate Screen", rot, 5),
                      ("draw scale 1.5 Normal", host.scale(f(1.5), f(1.5)), 0)]:
    report(name, timed(lambda: dev.draw(dst, src, m, mode)), n * n * 12)
report("draw scale 0.5 (minify + smooth)", timed(lambda: dev.draw(dst, src, host.scale(f(0.5), f(0.5)), 0)),
       n * n * 4 + n * n // 4 * 12)
report("drawTiled scale 0.37", timed(lambda: dev.draw_tiled(dst, src, host.scale(f(0.37), f(0.37)), 0)), n * n * 12)
ms = timed(lambda: dev.minify_by2(src, 1))
report("minifyBy2", ms, n * n * 5)
small = dev.minify_by2(src, 1)
report("magnifyBy2", timed(lambda: dev.magnify_by2(small, 1)), n * n * 5)
stops = [(0.0, (1, 0, 0, 1)), (0.3, (0, 1, 0, 0.5)), (1.0, (0, 0, 1, 1))]
for kind, handles, nm in [(3, [(10, 20), (n - 10, n - 30)], "linear"), (4, [(n / 2, n / 2), (n, n / 2), (n / 2, n)], "radial"),
                          (5, [(n / 2, n / 2), (n, n / 2), (n / 2, n)], "angular")]:
    report(f"fillGradient {nm}", timed(lambda: dev.fill_gradient(dst, kind, handles, stops, 1.0)), n * n * 4)
