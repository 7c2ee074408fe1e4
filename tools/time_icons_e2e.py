"""Icons end to end: host segment arrays -> cmdlist_create (+ H2D, count pass) -> run -> checksum, wall clock."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, synth  # noqa: E402
from pixie_b200.device import FillBatch  # noqa: E402

n_icons = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
size = 512
dev.init(0)
batch = FillBatch()
for i in range(n_icons):
    synth.icon_fills(i, size, i, batch)
arrays = batch.arrays()
img = dev.DeviceImage(size, size, n_icons)
tc, tr, tt = [], [], []
for it in range(6):
    img.fill(0)
    dev.sync()
    t0 = time.perf_counter()
    cl = dev.CmdList(size, size, n_icons, arrays)
    t1 = time.perf_counter()
    cl.run(img)
    c = img.checksum()
    t2 = time.perf_counter()
    del cl
    if it:
        tc.append(t1 - t0)
        tr.append(t2 - t1)
        tt.append(t2 - t0)
print("icons %d e2e: create %.3f ms  run+checksum %.3f ms  total %.3f ms  -> %.0f icons/s  (checksum %d)" % (
    n_icons, 1e3 * statistics.median(tc), 1e3 * statistics.median(tr), 1e3 * statistics.median(tt), n_icons / statistics.median(tt), c))
