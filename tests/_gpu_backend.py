"""Backend protocol of tests/golden_cases.py on the CUDA C ABI (pixie_cuda.so)."""
from __future__ import annotations

import numpy as np

from pixie_b200 import device as dev


class GpuBackend:
    """Every op: upload host pixels -> C-ABI call on device handles -> download."""

    def __init__(self):
        dev.init(0)
        self.covered = 0

    @staticmethod
    def _up(img):
        d = dev.DeviceImage(img.shape[1], img.shape[0], a8=(img.ndim == 2))
        d.upload(img)
        return d

    def fill_segments(self, img, segs, rgbx, rule, mode):
        d = self._up(img)
        b = dev.FillBatch()
        b.add(segs, rgbx, rule, mode)
        self.covered += dev.fill_batch(d, b.arrays(), count_covered=True)
        d.download(img)

    def blend_rect(self, dst, src, px, py, mode):
        d, s = self._up(dst), self._up(src)
        dev.blend_rect(d, s, px, py, mode)
        d.download(dst)

    def blend_rect_masked(self, dst, src, mask, px, py, mode):
        d, s, m = self._up(dst), self._up(src), self._up(mask)
        dev.blend_rect_masked(d, s, m, px, py, mode)
        d.download(dst)

    def blur(self, img, lut, radius, oob):
        d = self._up(img)
        dev.blur(d, lut, radius, oob)
        d.download(img)

    def spread(self, img, spread):
        d = self._up(img)
        dev.spread(d, spread)
        d.download(img)

    def shadow(self, img, ox, oy, spread, lut, radius, rgbx):
        s = self._up(img)
        d = dev.DeviceImage(img.shape[1], img.shape[0])
        dev.shadow(s, d, ox, oy, spread, lut, radius, rgbx)
        return d.download()

    def apply_opacity(self, img, opacity):
        d = self._up(img)
        dev.apply_opacity(d, opacity)
        d.download(img)

    def draw(self, dst, src, mat, mode):
        d, s = self._up(dst), self._up(np.ascontiguousarray(src))
        dev.draw(d, s, mat, mode)
        d.download(dst)

    def draw_tiled(self, dst, src, mat, mode, tiled=True):
        d, s = self._up(dst), self._up(np.ascontiguousarray(src))
        dev.draw_tiled(d, s, mat, mode, tiled)
        d.download(dst)

    def minify_by2(self, img, power=1):
        return dev.minify_by2(self._up(np.ascontiguousarray(img)), power).download()

    def magnify_by2(self, img, power=1):
        return dev.magnify_by2(self._up(np.ascontiguousarray(img)), power).download()

    def fill_gradient(self, img, kind, handles, stops, opacity=1.0):
        d = self._up(img)
        dev.fill_gradient(d, kind, handles, stops, opacity)
        d.download(img)
