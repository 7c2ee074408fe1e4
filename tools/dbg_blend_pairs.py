"""Finds pixels where the packed two-pixel float blend differs from the oracle; prints inputs and outputs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pixie_b200 import synth, common
from _gpu_backend import GpuBackend
from _oracle import OracleBackend
gb, ob = GpuBackend(), OracleBackend(0)
for mode in (8, 12, 13, 14, 15):
    tot = 0
    for seed in range(6):
        dst = synth.random_premultiplied(256, 1024, 11 + mode + 100 * seed)
        src = synth.random_premultiplied(256, 1024, 77 + mode + 100 * seed)
        a, b = dst.copy(), dst.copy()
        gb.blend_rect(a, src, 0, 0, mode)
        ob.blend_rect(b, src, 0, 0, mode)
        bad = np.argwhere((a != b).any(axis=2))
        tot += len(bad)
        for (y, x) in bad[:4]:
            print(common.BLEND_MODE_NAMES[mode], "dst", dst[y, x].tolist(), "src", src[y, x].tolist(), "gpu", a[y, x].tolist(), "oracle", b[y, x].tolist(), "x", int(x))
    print(common.BLEND_MODE_NAMES[mode], "differing px:", tot)
