"""Timeline of pixie_cuda_render_batch_host on the tiger (PIXIE_CUDA_TRACE=1 prints the band events)."""
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
pinned = dev.PinnedBuffer(size * size * 4)
for i in range(6):
    t0 = time.perf_counter()
    dev.render_batch_host(pinned.ptr, size, size, arrays)
    print("render_batch_host %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
