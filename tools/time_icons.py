"""BASELINE config 5 in miniature: 1024 synthetic 512^2 icons in one launch set, with the per-kernel split."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, synth  # noqa: E402
from pixie_b200.device import FillBatch  # noqa: E402

n_icons = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
size = 512
dev.init(0)
dev.set_profiling(True)
batch = FillBatch()
for i in range(n_icons):
    synth.icon_fills(i, size, i, batch)
arrays = batch.arrays()
img = dev.DeviceImage(size, size, n_icons)
cl = dev.CmdList(size, size, n_icons, arrays)
print(cl.info() if hasattr(cl, "info") else "")
ts, parts, plans, rasts = [], [], [], []
for it in range(6):
    img.fill(0)
    dev.timer_begin()
    cl.run(img)
    t = dev.timer_end()
    if it:
        ts.append(t)
        parts.append(dev.profile_read(dev.PROF_PARTITION))
        plans.append(dev.profile_read(dev.PROF_PLAN))
        rasts.append(dev.profile_read(dev.PROF_RASTER))
print("icons %d: step %.3f ms  partition %.3f  plan %.3f  raster %.3f" % (
    n_icons, statistics.median(ts), statistics.median(parts), statistics.median(plans), statistics.median(rasts)))
