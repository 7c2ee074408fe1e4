"""Helpers shared by the parity tests."""
from __future__ import annotations

import numpy as np

from pixie_b200.host import Segments


def oracle_render_batch(arrays, width, height, layers=1, background=None, sem=0):
    """Apply an ordered fill command list with the CPU oracle, one fill at a time."""
    from _oracle import OracleBackend

    ob = OracleBackend(sem)
    canv = np.zeros((layers, height, width, 4), np.uint8)
    if background is not None:
        canv[...] = background
    for k in range(len(arrays["rgbx"])):
        s0, s1 = int(arrays["seg_offsets"][k]), int(arrays["seg_offsets"][k + 1])
        segs = Segments(arrays["xyxy"][s0:s1], arrays["winding"][s0:s1])
        ob.fill_segments(canv[int(arrays["layer"][k])], segs, int(arrays["rgbx"][k]), int(arrays["rule"][k]),
                         int(arrays["mode"][k]))
    return canv, ob.covered


def gpu_render_batch(arrays, width, height, layers=1, background=None):
    from pixie_b200 import device as dev

    dev.init(0)
    img = dev.DeviceImage(width, height, layers)
    if background is not None:
        canv = np.zeros((layers, height, width, 4), np.uint8)
        canv[...] = background
        img.upload(canv)
    covered = dev.fill_batch(img, arrays, count_covered=True)
    return img.download().reshape(layers, height, width, 4), covered


def diff_report(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    bad = d.max(axis=-1) > 0
    where = np.argwhere(bad)[:5].tolist()
    return int(bad.sum()), int(d.max()), where
