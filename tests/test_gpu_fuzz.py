"""Differential fuzzing of the CUDA rasteriser against the oracle on raw segment soups — inputs the
shape producers never make: unbalanced winding, coincident and overlapping edges, segments far outside
the canvas, sub-pixel slivers, integer-aligned verticals mixed with slanted edges, >2 hits at one x."""
import os

import numpy as np
import pytest

from pixie_b200.device import FillBatch
from pixie_b200.host import Segments
from _util import diff_report, gpu_render_batch, oracle_render_batch

pytestmark = pytest.mark.gpu


def _soup(rng, n, w, h, style):
    if style == "grid":        # quarter-pixel lattice: many exact coincidences and pixel-aligned edges
        x = rng.integers(-8, 4 * w + 8, (n, 2)) / 4.0
        y = np.sort(rng.integers(-8, 4 * h + 8, (n, 2)) / 4.0, axis=1)
    elif style == "wide":      # far outside the canvas on every side
        x = rng.uniform(-3 * w, 4 * w, (n, 2))
        y = np.sort(rng.uniform(-3 * h, 4 * h, (n, 2)), axis=1)
    elif style == "sliver":    # nearly vertical / nearly horizontal, short
        cx, cy = rng.uniform(0, w, n), rng.uniform(0, h, n)
        dx, dy = rng.normal(0, 0.7, n), np.abs(rng.normal(0, 3.0, n)) + 1 / 256
        x = np.stack([cx, cx + dx], 1)
        y = np.stack([cy, cy + dy], 1)
    else:                      # uniform
        x = rng.uniform(-4, w + 4, (n, 2))
        y = np.sort(rng.uniform(-4, h + 4, (n, 2)), axis=1)
    y = np.floor(y * 256) / 256                      # shapesToSegments quantises y to 1/256
    keep = y[:, 0] < y[:, 1]                         # ... and drops horizontals; at.y < to.y
    xy = np.stack([x[:, 0], y[:, 0], x[:, 1], y[:, 1]], 1)[keep].astype(np.float32)
    if style == "grid":        # duplicate some edges with opposite / equal winding
        dup = xy[rng.integers(0, len(xy), len(xy) // 3)]
        xy = np.concatenate([xy, dup])
    wind = rng.choice(np.array([1, -1], np.int16), len(xy))
    return Segments(xy, wind)


@pytest.mark.parametrize("style", ["uniform", "grid", "wide", "sliver"])
@pytest.mark.parametrize("seed", range(6))
def test_segment_soup(style, seed):
    rng = np.random.default_rng([seed, ["uniform", "grid", "wide", "sliver"].index(style)])
    w, h = [(96, 64), (64, 96), (130, 37), (33, 33), (256, 16), (20, 200)][seed]
    b = FillBatch()
    for k in range(6):
        n = int(rng.integers(2, 120))
        segs = _soup(rng, n, w, h, style)
        if len(segs) == 0:
            continue
        col = int(rng.integers(0, 2 ** 32)) if k % 2 else 0xFF000000 | int(rng.integers(0, 2 ** 24))
        mode = [0, 17, 16, 0, 11, 19][k]
        b.add(segs, col, int(rng.integers(0, 2)), mode)
    arr = b.arrays()
    bg = rng.integers(0, 256, (1, h, w, 4), dtype=np.uint8)
    want, wc = oracle_render_batch(arr, w, h, background=bg)
    got, gc_ = gpu_render_batch(arr, w, h, background=bg)
    n, mx, where = diff_report(got, want)
    if n and os.path.isdir("gpurun_out"):
        np.savez(f"gpurun_out/fuzz_fail_{style}_{seed}.npz", got=got, want=want, bg=bg, **arr)
    assert n == 0, f"{style}/{seed}: {n} px differ (max {mx}) at {where}"
    assert gc_ == wc


@pytest.mark.parametrize("style", ["uniform", "grid", "wide", "sliver"])
@pytest.mark.parametrize("w", [6144, 4100, 4101, 9001])
def test_segment_soup_wide_canvas_row_tiles(style, w):
    """Canvases wider than one raster tile (2048 columns): rows are rasterised by several warps, each clamped to
    its tile — spans, trapezoid edges and MaskBlend clears that cross tile boundaries, vector and scalar rows."""
    h = 24
    rng = np.random.default_rng([w, ["uniform", "grid", "wide", "sliver"].index(style)])
    b = FillBatch()
    for k in range(8):
        n = int(rng.integers(2, 160))
        segs = _soup(rng, n, w, h, style)
        if len(segs) == 0:
            continue
        col = int(rng.integers(0, 2 ** 32)) if k % 2 else 0xFF000000 | int(rng.integers(0, 2 ** 24))
        mode = [0, 17, 16, 0, 11, 19, 0, 16][k]
        b.add(segs, col, int(rng.integers(0, 2)), mode)
    # axis-aligned rectangles that straddle the tile boundaries (mode A / trapezoid shortcuts)
    for x0, x1 in ((2040.0, 2060.0), (1000.5, 5000.25), (2048.0, 4096.0), (-10.0, w + 10.0)):
        xy = np.array([[x0, 2, x0, 20], [x1, 2, x1, 20]], np.float32)
        b.add(Segments(xy, np.array([1, -1], np.int16)), 0x80402010, 0, int(rng.choice([0, 16, 17])))
    arr = b.arrays()
    bg = rng.integers(0, 256, (1, h, w, 4), dtype=np.uint8)
    want, wc = oracle_render_batch(arr, w, h, background=bg)
    got, gc_ = gpu_render_batch(arr, w, h, background=bg)
    n, mx, where = diff_report(got, want)
    assert n == 0, f"{style}/{w}: {n} px differ (max {mx}) at {where}"
    assert gc_ == wc
