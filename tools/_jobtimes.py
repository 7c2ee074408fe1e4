import os, sys
sys.path.insert(0, '/root/repo')
os.environ["PIXIE_CUDA_JOBTIMES"] = "gpurun_out/jobtimes.bin"
import numpy as np
from bench import tiger_arrays
from pixie_b200 import device as dev
dev.init(0)
arrays = tiger_arrays(4096)
img = dev.DeviceImage(4096, 4096)
cl = dev.CmdList(4096, 4096, 1, arrays)
for _ in range(3):
    img.fill(0); cl.run(img)
dev.sync()
a = np.fromfile("gpurun_out/jobtimes.bin", dtype=np.int64).reshape(-1, 4)
a = a[a[:, 0] != 0]
t0 = a[:, 0].min()
dur = a[:, 1] - a[:, 0]
ecnt = a[:, 2] >> 32
kn = a[:, 2] & 0xFFFFFFFF
print("heavy jobs", len(a), "span cycles", a[:, 1].max() - t0)
order = np.argsort(-dur)[:15]
for i in order:
    print("dur %8d start %8d eCnt %4d kind %d n %5d warp %x" % (dur[i], a[i, 0] - t0, ecnt[i], kn[i] // 100000, kn[i] % 100000, a[i, 3]))
print("sum dur", dur.sum(), "mean", dur.mean())
for lo, hi in ((17, 32), (33, 64), (65, 128), (129, 256), (257, 1024)):
    m = (ecnt >= lo) & (ecnt <= hi)
    if m.any(): print("eCnt %d-%d: %d jobs, mean dur %.0f, max %d, last end %d" % (lo, hi, m.sum(), dur[m].mean(), dur[m].max(), (a[m, 1] - t0).max()))
# per-warp finish
