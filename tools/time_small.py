"""Latency of small calls through the C ABI: examples/heart.nim (one fill on a 200x200 canvas) on device handles and
through the host-pointer variant, next to the CPU oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_cases as gc  # noqa: E402
from pixie_b200 import device as dev, host  # noqa: E402

dev.init(0)
segs = host.fill_segments(gc.HEART, None)
rgbx = gc.html("#FC427B")
img = dev.DeviceImage(200, 200)
host_px = np.full((200, 200, 4), 255, np.uint8)


def t(fn, n=200):
    for _ in range(10):
        fn()
    dev.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dev.sync()
    return (time.perf_counter() - t0) / n * 1e6


print("fill_segments on a device handle (async, amortised)  %.1f us" % t(lambda: dev.fill_segments(img, segs, rgbx, 0, 0)))
print("fill_segments + sync (latency)                       %.1f us" % t(lambda: (dev.fill_segments(img, segs, rgbx, 0, 0), dev.sync())))
L = dev.lib()
print("fill_segments_host (upload + fill + download)        %.1f us" % t(lambda: dev.check(L.pixie_cuda_fill_segments_host(
    host_px.ctypes.data, 200, 200, segs.xyxy.ctypes.data, segs.winding.ctypes.data, len(segs), rgbx, 0, 0))))
from _oracle import OracleBackend  # noqa: E402

ob = OracleBackend(0)
cpu = host_px.copy()
t0 = time.perf_counter()
for _ in range(200):
    ob.fill_segments(cpu, segs, rgbx, 0, 0)
print("CPU oracle, same fill                                %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
