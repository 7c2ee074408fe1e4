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
a = np.fromfile("gpurun_out/jobtimes.bin", dtype=np.int64).reshape(-1, 16)
a = a[a[:, 0] != 0]
ecnt = a[:, 15] >> 32; nsel = a[:, 15] & 0xFFFFFFFF
nst = (a[:, :14] != 0).sum(1)
last = a[np.arange(len(a)), nst - 1]
dur = last - a[:, 0]
for i in np.argsort(-dur)[:8]:
    st = a[i, :nst[i]] - a[i, 0]
    print("eCnt %d nsel %d dur %d phases %s" % (ecnt[i], nsel[i], dur[i], np.diff(st).tolist()))
