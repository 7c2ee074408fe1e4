"""The window method of tests/_windows.py, checked on the CPU: the oracle run on a crop (window + reach) reproduces
the window of the oracle's whole-image result — for blur, spread and shadow, at corners, borders and inside."""
import numpy as np

from pixie_b200 import host, synth
from pixie_b200.common import rgbx as pack
from _oracle import OracleBackend
import _windows as W


class _HostImage:
    """Stands in for a DeviceImage: download_rows over a host array (the whole-image oracle result)."""

    def __init__(self, a):
        self.a = a

    def download_rows(self, y0, y1):
        return self.a[y0:y1]


def _windows(h, w):
    return W.corner_and_seam_windows(h, w, size=24, seams=((h // 2, w // 3), (h // 3 + 7, w // 2 + 5)))


def test_blur_windows_equal_whole_image():
    h, w, r = 150, 210, 9
    img = synth.random_premultiplied(h, w, 3)
    lut = host.gaussianKernel(r)
    for oob in (0, pack(10, 20, 30, 200)):
        whole = img.copy()
        OracleBackend(0).blur(whole, lut, r, oob)
        n, bad, mx = W.check_blur_windows(_HostImage(whole), img, lut, r, oob, _windows(h, w))
        assert n > 0 and bad == 0, (bad, mx)


def test_spread_windows_equal_whole_image():
    h, w = 120, 170
    img = synth.random_premultiplied(h, w, 4)
    for amount in (3, -2):
        whole = img.copy()
        OracleBackend(0).spread(whole, amount)
        n, bad, mx = W.check_spread_windows(_HostImage(whole), img, amount, _windows(h, w))
        assert n > 0 and bad == 0, (bad, mx)


def test_shadow_windows_equal_whole_image():
    h, w, r = 140, 190, 6
    img = synth.random_premultiplied(h, w, 5)
    img[: h // 4] = 0
    lut = host.gaussianKernel(r)
    for offset in ((3, -2), (8, 8)):
        whole = OracleBackend(0).shadow(img, offset[0], offset[1], 2, lut, r, pack(0, 0, 0, 200))
        n, bad, mx = W.check_shadow_windows(_HostImage(whole), img, offset, 2, lut, r, pack(0, 0, 0, 200), _windows(h, w))
        assert n > 0 and bad == 0, (bad, mx)


def test_window_check_detects_a_wrong_pixel():
    h, w, r = 64, 64, 4
    img = synth.random_premultiplied(h, w, 6)
    lut = host.gaussianKernel(r)
    whole = img.copy()
    OracleBackend(0).blur(whole, lut, r, 0)
    whole[2, 3, 1] ^= 1
    n, bad, mx = W.check_blur_windows(_HostImage(whole), img, lut, r, 0, [(0, 16, 0, 16)])
    assert bad == 1 and mx == 1
