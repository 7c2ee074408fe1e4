"""Deterministic synthetic workloads: icon/glyph-like path sets,
premultiplied RGBX noise images and AA coverage masks.  Host-side input generation only."""
from __future__ import annotations

import math

import numpy as np

from . import host
from .common import NormalBlend, OverwriteBlend, rgba_to_rgbx
from .device import FillBatch


def _blob(rng, cx, cy, r, knots):
    """Random closed cubic blob with `knots` knots around (cx, cy)."""
    p = host.newPath()
    ang = np.sort(rng.uniform(0, 2 * math.pi, knots))
    rad = rng.uniform(0.45 * r, r, knots)
    pts = [(cx + rad[i] * math.cos(ang[i]), cy + rad[i] * math.sin(ang[i])) for i in range(knots)]
    p.moveTo(*pts[0])
    for i in range(knots):
        a, b = pts[i], pts[(i + 1) % knots]
        j = rng.uniform(-0.35 * r, 0.35 * r, 4)
        p.bezierCurveTo(a[0] + (b[0] - a[0]) / 3 + j[0], a[1] + (b[1] - a[1]) / 3 + j[1],
                        a[0] + 2 * (b[0] - a[0]) / 3 + j[2], a[1] + 2 * (b[1] - a[1]) / 3 + j[3], b[0], b[1])
    p.closePath()
    return p


def icon_fills(index: int, size: int = 512, layer: int = 0, batch: FillBatch | None = None, seed: int = 0, paths=None):
    """One synthetic icon: 1-8 sub-paths of rect / roundedRect / ellipse / polygon /
    closed cubic blobs / nested even-odd contours, optional strokes; seed = icon index; transparent
    canvas, first fill OverwriteBlend then NormalBlend (svg.nim:562,590).  With `paths` (a device.PathBatch) the icon is
    appended as path commands instead of host-flattened segments."""
    rng = np.random.default_rng([seed, index])
    b = batch if batch is not None else FillBatch()
    lo, hi = size / 32.0, size - size / 32.0
    first = True
    for _ in range(int(rng.integers(1, 9))):
        kind = int(rng.integers(0, 6))
        cx, cy = rng.uniform(lo + size / 8, hi - size / 8, 2)
        r = float(rng.uniform(size / 16, size / 3))
        p = host.newPath()
        rule = int(rng.integers(0, 2))
        if kind == 0:
            w, h = rng.uniform(size / 16, size / 2, 2)
            if rng.random() < 0.5:  # pixel-aligned rect exercises the no-AA modes
                p.rect(float(int(cx - w / 2)), float(int(cy - h / 2)), float(int(w)), float(int(h)))
            else:
                p.rect(cx - w / 2, cy - h / 2, w, h)
        elif kind == 1:
            w, h = rng.uniform(size / 8, size / 2, 2)
            rr = rng.uniform(0, min(w, h) / 2, 4)
            p.roundedRect(cx - w / 2, cy - h / 2, w, h, *[float(v) for v in rr])
        elif kind == 2:
            p.ellipse(cx, cy, r, float(rng.uniform(size / 16, size / 3)))
        elif kind == 3:
            p.polygon(cx, cy, r, int(rng.integers(3, 9)))
        elif kind == 4:
            p = _blob(rng, cx, cy, r, int(rng.integers(4, 13)))
        else:  # glyph-like: two nested contours, hole through even-odd
            p.ellipse(cx, cy, r, r * 0.8)
            p.ellipse(cx, cy, r * 0.55, r * 0.45)
            rule = host.EvenOdd
        col = [int(v) for v in rng.integers(0, 256, 3)]
        alpha = 255 if rng.random() < 0.5 else 153
        rgbx = rgba_to_rgbx(col[0], col[1], col[2], alpha)
        tr = None
        if rng.random() < 0.3:
            a = float(rng.uniform(-0.5, 0.5))
            tr = host.matmul(host.translate(cx, cy), host.matmul(host.rotate(a), host.translate(-cx, -cy)))
        mode = OverwriteBlend if first else NormalBlend
        if rng.random() < 0.3:
            sw = float(rng.uniform(1, size * 24 / 512))
            cap, join = int(rng.integers(0, 3)), int(rng.integers(0, 3))
            if paths is not None:  # the same icon as path COMMANDS (flattened and stroked on the device)
                paths.add_stroke(p, tr, sw, cap, join, host.defaultMiterLimit, (), rgbx, host.NonZero, mode, layer)
            else:
                b.add(host.stroke_segments(p, tr, sw, cap, join), rgbx, host.NonZero, mode, layer)
        elif paths is not None:
            paths.add_fill(p, tr, rgbx, rule, mode, layer)
        else:
            b.add(host.fill_segments(p, tr), rgbx, rule, mode, layer)
        first = False
    return b if paths is None else paths


def random_premultiplied(h: int, w: int, seed: int) -> np.ndarray:
    """RGBX premultiplied noise: a ~ U{0..255}, c ~ U{0..a}; 64-px runs with 5 % a=0
    and 20 % a=255."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w), dtype=np.uint16)
    runs = rng.random((h, (w + 63) // 64))
    sel = np.repeat(runs, 64, axis=1)[:, :w]
    a[sel < 0.05] = 0
    a[sel > 0.80] = 255
    c = (rng.integers(0, 256, (h, w, 3), dtype=np.uint16) * (a[..., None] + 1)) >> 8
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = np.minimum(c, a[..., None])
    out[..., 3] = a
    return out


def coverage_mask(h: int, w: int, seed: int) -> np.ndarray:
    """8-bit AA coverage plane: 64-px runs, 30 % 0, 50 % 255, 20 % partial."""
    rng = np.random.default_rng(seed)
    runs = rng.random((h, (w + 63) // 64))
    sel = np.repeat(runs, 64, axis=1)[:, :w]
    m = rng.integers(1, 255, (h, w), dtype=np.uint8)
    m[sel < 0.30] = 0
    m[sel > 0.50] = 255
    return m
