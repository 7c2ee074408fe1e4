import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from pixie_b200 import device as dev, host
from _oracle import OracleBackend
dev.init(0)
n = int(sys.argv[1])
rng = np.random.default_rng(1)
px = rng.integers(0, 256, (n, n, 4), dtype=np.uint8)
src = dev.DeviceImage(n, n); dst = dev.DeviceImage(n, n)
src.upload(px)
R = 32 if len(sys.argv) > 2 else 0
lut = host.gaussianKernel(R)
print(lut)
for i in range(3):
    dev.sync(); t0 = time.perf_counter()
    dev.shadow(src, dst, 8, 8, 4, lut, R, 0xC8000000)
    dev.sync(); print((time.perf_counter() - t0) * 1e3, "ms")
if len(sys.argv) > 2: sys.exit(0)
out = dst.download()
ob = OracleBackend(0)
ref = ob.shadow(px, 8, 8, 4, lut, 0, 0xC8000000)
print("diff", int((out != ref).sum()), out[100, 100], ref[100, 100])
