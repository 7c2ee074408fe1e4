"""Breaks the end-to-end tiger step into its parts (wall clock, synchronised)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import tiger_arrays  # noqa: E402
from pixie_b200 import device as dev  # noqa: E402

dev.init(0)
size = 4096
arrays = tiger_arrays(size)
img = dev.DeviceImage(size, size)
pinned = dev.PinnedBuffer(size * size * 4)


def t(fn, n=10):
    fn()
    dev.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        dev.sync()
    return (time.perf_counter() - t0) / n * 1e3


print("clear            %.3f ms" % t(lambda: img.fill(0)))
print("fill_batch       %.3f ms (host precompute + H2D + partition + raster)" % t(lambda: dev.fill_batch(img, arrays)))
cl = dev.CmdList(size, size, 1, arrays)
print("cmdlist.run      %.3f ms (partition + raster)" % t(lambda: cl.run(img)))
print("download pinned  %.3f ms -> %.1f GB/s" % ((lambda ms: (ms, size * size * 4 / ms / 1e6))(t(lambda: dev.download_async(img, pinned)))))
import numpy as np
host = np.empty((size, size, 4), np.uint8)
print("download pageable %.3f ms" % t(lambda: img.download(host)))
