"""Quick device timings of the blur / blend kernels (CUDA events via the library's profiling slots)."""
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, host, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev.init(0)
dev.set_profiling(True)
if which in ("blur", "all"):
    for n, r in ((16384, 32), (16384, 8), (8192, 64), (4096, 3)):
        img = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(256, n, 4), (n // 256, 1, 1)))
        lut = host.gaussianKernel(r)
        xs, ys = [], []
        for it in range(4):
            dev.blur(img, lut, r, 0)
            if it:
                xs.append(dev.profile_read(dev.PROF_BLUR_X))
                ys.append(dev.profile_read(dev.PROF_BLUR_Y))
        x, y = statistics.median(xs), statistics.median(ys)
        print(f"blur {n}^2 r={r}: X {x:.3f} ms  Y {y:.3f} ms  total {x + y:.3f} ms  -> {n * n * 8 / (x + y) / 1e6:.0f} GB/s (8 B/px), "
              f"{n * n * 4 * 2 * (2 * r + 1) / (x + y) / 1e9:.2f} Ttap/s")
        del img
if which in ("blend", "all"):
    n = 8192
    dst0 = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 1), (n // 512, 1, 1)))
    dst = dev.DeviceImage(n, n)
    src = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 2), (n // 512, 1, 1)))
    mask = dev.DeviceImage(n, n, a8=True).upload(np.tile(synth.coverage_mask(512, n, 3), (n // 512, 1)))
    for masked in (True, False):
        for mode in (0, 16, 17, 2, 4, 7, 3, 8):
            ts = []
            for it in range(4):
                dst.copy_from(dst0)
                if masked:
                    dev.blend_rect_masked(dst, src, mask, 0, 0, mode)
                else:
                    dev.blend_rect(dst, src, 0, 0, mode)
                if it:
                    ts.append(dev.profile_read(dev.PROF_BLEND))
            t = statistics.median(ts)
            bpp = (8 if mode == 17 else 12) + (1 if masked else 0)
            print(f"blend mode {mode:2d} masked={int(masked)}: {t * 1e3:.1f} us  {n * n * bpp / t / 1e6:.0f} GB/s ({bpp} B/px)")
