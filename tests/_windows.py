"""Oracle parity at BASELINE sizes through windows (test infrastructure: uses the CPU oracle as the checker).

A 16384^2 blur takes the scalar oracle many minutes, but the filters are local: output pixel (y, x) of
`blur` depends on the input rows [y - r, y + r] x columns [x - r, x + r] only (two separable passes with an 8-bit
intermediate that is itself a function of one row, images.nim:304-365), `spread` likewise with |spread|, and `shadow`
(offset copy -> spread -> blur -> MaskBlend composite, images.nim:760-776) with ceil|offset| + |spread| + radius.  So
the oracle run on a CROP that extends the window by that reach reproduces the window of the full-size result exactly:
where the crop is cut inside the image the cut is at least `reach` away from the window, and where it ends at the
image border the out-of-bounds rule is the image's own.  Blends are per pixel: any row range is a window.
"""
from __future__ import annotations

import math

import numpy as np

from _oracle import OracleBackend
from _util import diff_report


def corner_and_seam_windows(h, w, size=48, seams=((8192, 4096), (4096 + 128, 8192 + 32))):
    """Four image corners, windows across kernel tile seams (multiples of 128 / 32), one interior."""
    s = min(size, h, w)
    wins = [(0, s, 0, s), (0, s, w - s, w), (h - s, h, 0, s), (h - s, h, w - s, w)]
    for (sy, sx) in seams:
        sy, sx = min(sy, h - s) - s // 2, min(sx, w - s) - s // 2
        wins.append((max(0, sy), max(0, sy) + s, max(0, sx), max(0, sx) + s))
    cy, cx = (h * 5) // 11, (w * 7) // 13
    wins.append((cy, min(h, cy + s), cx, min(w, cx + s)))
    return wins


def _crop(inp, win, reach):
    y0, y1, x0, x1 = win
    h, w = inp.shape[:2]
    cy0, cy1, cx0, cx1 = max(0, y0 - reach), min(h, y1 + reach), max(0, x0 - reach), min(w, x1 + reach)
    return inp[cy0:cy1, cx0:cx1].copy(), (y0 - cy0, y1 - cy0, x0 - cx0, x1 - cx0)


def _gpu_window(dimg, win):
    y0, y1, x0, x1 = win
    return dimg.download_rows(y0, y1)[:, x0:x1]


def check_blur_windows(dimg, inp, lut, radius, oob, windows):
    """dimg: DeviceImage holding blur(inp); returns (pixels compared, mismatching pixels, max |delta|)."""
    ob = OracleBackend(0)
    n = bad = mx = 0
    for win in windows:
        crop, (a, b, c, d) = _crop(inp, win, radius)
        ob.blur(crop, lut, radius, oob)
        nb, m, _ = diff_report(_gpu_window(dimg, win), crop[a:b, c:d])
        n += (b - a) * (d - c)
        bad += nb
        mx = max(mx, m)
    return n, bad, mx


def check_spread_windows(dimg, inp, amount, windows):
    ob = OracleBackend(0)
    n = bad = mx = 0
    for win in windows:
        crop, (a, b, c, d) = _crop(inp, win, abs(amount))
        ob.spread(crop, amount)
        nb, m, _ = diff_report(_gpu_window(dimg, win), crop[a:b, c:d])
        n += (b - a) * (d - c)
        bad += nb
        mx = max(mx, m)
    return n, bad, mx


def check_shadow_windows(dimg, inp, offset, spread, lut, radius, rgbx, windows):
    ob = OracleBackend(0)
    reach = int(math.ceil(max(abs(offset[0]), abs(offset[1])))) + abs(spread) + radius
    n = bad = mx = 0
    for win in windows:
        crop, (a, b, c, d) = _crop(inp, win, reach)
        out = ob.shadow(crop, offset[0], offset[1], spread, lut, radius, rgbx)
        nb, m, _ = diff_report(_gpu_window(dimg, win), out[a:b, c:d])
        n += (b - a) * (d - c)
        bad += nb
        mx = max(mx, m)
    return n, bad, mx


def check_blend_rows(ddst, dst_in, src, mask, mode, row_ranges):
    """ddst: DeviceImage holding blend_rect_masked(dst_in, src, mask, 0, 0, mode) (mask None: blend_rect)."""
    ob = OracleBackend(0)
    n = bad = mx = 0
    for (y0, y1) in row_ranges:
        want = dst_in[y0:y1].copy()  # a slice of a contiguous array IS contiguous: ascontiguousarray would alias it
        s = np.ascontiguousarray(src[y0:y1])
        if mask is None:
            ob.blend_rect(want, s, 0, 0, mode)
        else:
            ob.blend_rect_masked(want, s, np.ascontiguousarray(mask[y0:y1]), 0, 0, mode)
        nb, m, _ = diff_report(ddst.download_rows(y0, y1), want)
        n += want.shape[0] * want.shape[1]
        bad += nb
        mx = max(mx, m)
    return n, bad, mx
