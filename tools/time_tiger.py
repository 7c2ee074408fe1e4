"""Device time of one tiger step (clear + K1..K3) and of its raster kernel; PIXIE_CUDA_LIB=<other build> for A/B runs."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import tiger_arrays  # noqa: E402
from pixie_b200 import device as dev  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev.init(0)
dev.set_profiling(True)
arrays = tiger_arrays(size)
img = dev.DeviceImage(size, size)
cl = dev.CmdList(size, size, 1, arrays)
steps, rast, plan, part = [], [], [], []
for it in range(25):
    dev.timer_begin()
    if os.environ.get("TIGER_CLEAR"):
        cl.run(img, clear=True)  # the raster kernel clears the canvas
    else:
        img.fill(0)
        cl.run(img)
    t = dev.timer_end()
    if it >= 5:
        steps.append(t)
        try:
            rast.append(dev.profile_read(dev.PROF_RASTER))
            plan.append(dev.profile_read(dev.PROF_PLAN))
            part.append(dev.profile_read(dev.PROF_PARTITION))
        except Exception:
            rast.append(float('nan'))
print("tiger %d^2 [%s]: step %.4f ms (min %.4f)  partition %.4f  plan %.4f  raster %.4f ms" % (
    size, os.path.basename(os.environ.get("PIXIE_CUDA_LIB", "in-tree")), statistics.median(steps), min(steps),
    statistics.median(part) if part else float("nan"), statistics.median(plan) if plan else float("nan"), statistics.median(rast)))
