"""Parity of the CUDA rasteriser (through pixie_cuda_fill_batch) with the CPU oracle: bit-exact."""
import numpy as np
import pytest

from pixie_b200 import host, synth
from pixie_b200.common import NormalBlend, OverwriteBlend, MaskBlend, rgbx as pack
from pixie_b200.device import FillBatch
from _util import diff_report, gpu_render_batch, oracle_render_batch

pytestmark = pytest.mark.gpu


def _check(arrays, w, h, layers=1, background=None):
    want, wc = oracle_render_batch(arrays, w, h, layers, background)
    got, gc_ = gpu_render_batch(arrays, w, h, layers, background)
    n, mx, where = diff_report(got, want)
    assert n == 0, f"{n} px differ (max {mx}) at {where}"
    assert gc_ == wc, f"covered {gc_} != {wc}"


@pytest.mark.parametrize("seed", range(12))
def test_icons_svg_semantics(seed):
    """Synthetic icons (C5 generator), first fill Overwrite then Normal, 4 layers in one batch."""
    b = FillBatch()
    for layer in range(4):
        synth.icon_fills(seed * 4 + layer, 256, layer, b)
    _check(b.arrays(), 256, 256, layers=4)


@pytest.mark.parametrize("mode", range(20))
def test_every_blend_mode_on_fills(mode):
    """All 20 BlendMode enumerators as the fill's blend mode over a noisy premultiplied backdrop."""
    bg = synth.random_premultiplied(128, 128, 100 + mode)
    b = FillBatch()
    for i in range(3):
        synth.icon_fills(1000 + 7 * mode + i, 128, 0, b)
    arr = b.arrays()
    arr["mode"][:] = mode
    _check(arr, 128, 128, background=bg)


@pytest.mark.parametrize("size", [(61, 47), (130, 33), (257, 19), (4, 4), (1, 1), (7, 300)])
def test_odd_canvas_sizes(size):
    """Widths that are not a multiple of 4 take the scalar row path; shapes hang over every edge."""
    w, h = size
    rng = np.random.default_rng(w * 1000 + h)
    b = FillBatch()
    for i in range(10):
        p = host.newPath()
        cx, cy = rng.uniform(-0.3 * w, 1.3 * w), rng.uniform(-0.3 * h, 1.3 * h)
        kind = i % 3
        if kind == 0:
            p.ellipse(cx, cy, rng.uniform(1, w), rng.uniform(1, h))
        elif kind == 1:
            p.rect(float(int(cx)), float(int(cy)), float(int(rng.uniform(1, w))), float(int(rng.uniform(1, h))))
        else:
            p.polygon(cx, cy, rng.uniform(1, max(w, h)), int(rng.integers(3, 9)))
        col = pack(*[int(v) for v in rng.integers(0, 256, 3)], 255) if i % 2 else pack(40, 30, 20, 90)
        mode = [NormalBlend, OverwriteBlend, MaskBlend, 11, 19][i % 5]
        if i % 4 == 3:
            b.add(host.stroke_segments(p, None, float(rng.uniform(0.5, 9))), col, host.NonZero, mode)
        else:
            b.add(host.fill_segments(p), col, int(rng.integers(0, 2)), mode)
    bg = synth.random_premultiplied(h, w, 5)
    _check(b.arrays(), w, h, background=bg)


def test_mask_blend_clears_outside_and_empty_paths():
    """MaskBlend clears everything the path does not cover (paths.nim:1516-1517,1564-1582,1856-1869,
    1910-1912); zero-width paths leave the image untouched (:1615-1616); empty paths draw nothing."""
    w = h = 96
    bg = synth.random_premultiplied(h, w, 9)
    cases = [
        "M 20.5 10.5 L 70.5 30.5 L 40.5 80.5 z",          # AA coverage rows + trapezoid rows
        "M 10 10 H 60 V 60 H 10 z",                        # pixel aligned (mode A)
        "M 200 200 H 260 V 260 H 200 z",                   # entirely off canvas: pathWidth == 0 -> untouched
        "M -50 20 H 30 V 60 H -50 z",                      # hangs over the left edge
        "M 0 0 L 0 1 L 0 0 Z",                             # zero width
        "M 5 5 z",                                         # empty
        "M 10 -40 L 90 -20 L 50 -5 z",                     # above the canvas: MaskBlend clears all rows
    ]
    for d in cases:
        b = FillBatch()
        b.add(host.fill_segments(d), pack(0, 255, 0, 255), host.NonZero, MaskBlend)
        b.add(host.fill_segments(d), pack(0, 100, 0, 100), host.EvenOdd, MaskBlend)
        _check(b.arrays(), w, h, background=bg)


def test_many_entries_per_band_spills_to_global_scratch():
    """A stroke made of hundreds of overlapping polygons puts far more than 64 entries in a band."""
    rng = np.random.default_rng(3)
    p = host.newPath()
    p.moveTo(20, 100)
    for i in range(400):
        p.lineTo(20 + i * 0.9, 100 + 60 * np.sin(i * 0.7) + rng.uniform(-3, 3))
    b = FillBatch()
    b.add(host.stroke_segments(p, None, 7.0, host.RoundCap, host.RoundJoin), pack(200, 10, 10, 255), host.NonZero,
          NormalBlend)
    b.add(host.stroke_segments(p, None, 3.0), pack(0, 0, 80, 128), host.NonZero, NormalBlend)
    _check(b.arrays(), 400, 200)


def test_single_fill_entry_point_and_host_variant():
    """pixie_cuda_fill_segments and pixie_cuda_fill_segments_host give the same pixels as the batch."""
    import ctypes as C
    from pixie_b200 import device as dev
    from _oracle import OracleBackend

    dev.init(0)
    segs = host.fill_segments("M 10.2 5 C 80 -20 120 100 40 90 S 0 40 10.2 5 z", host.scale(1.5, 1.5))
    want = np.full((150, 200, 4), 255, np.uint8)
    OracleBackend(0).fill_segments(want, segs, pack(10, 20, 30, 200), host.NonZero, NormalBlend)
    img = dev.DeviceImage(200, 150)
    img.fill(0xFFFFFFFF)
    dev.fill_segments(img, segs, pack(10, 20, 30, 200), host.NonZero, NormalBlend)
    assert diff_report(img.download(), want)[0] == 0
    px = np.full((150, 200, 4), 255, np.uint8)
    dev.check(dev.lib().pixie_cuda_fill_segments_host(px.ctypes.data, 200, 150, segs.xyxy.ctypes.data,
                                                      segs.winding.ctypes.data, len(segs), pack(10, 20, 30, 200), 0, 0))
    assert diff_report(px, want)[0] == 0


@pytest.mark.parametrize("pairs", [((-1.45, -1.2), (-0.5, -0.2), (0.25, 0.6)), ((-3.4, -2.9), (-1.7, -1.2), (0.25, 0.6)),
                                   ((-0.9, -0.1), (0.25, 0.6)), ((-5.0, -4.5), (-0.3, 0.3), (0.5, 1.4))])
@pytest.mark.parametrize("width", [40, 37])
def test_mask_clears_left_of_canvas_wrap_into_rows_above(width, pairs):
    """MaskBlend through the trapezoid shortcut with fill pairs left of x = 0: the reference's
    clearUnsafe(min(filledTo, w), y, min(clearTo, w), y) (paths.nim:1856-1866, :1433-1440) addresses the
    canvas linearly, so a negative x range clears pixels of the rows above — inside the image, hence
    part of the reference's result (the oracle does the same; only indices < 0 are dropped)."""
    w, h = width, 24
    rows, wind = [], []
    for x0, x1 in pairs:
        rows.append([x0 * w, -2.0, x0 * w + 3.0, h + 2.0])
        rows.append([x1 * w + 2.0, -2.0, x1 * w, h + 2.0])
        wind += [1, -1]
    segs = host.Segments(np.array(rows, np.float32), np.array(wind, np.int16))
    bg = synth.random_premultiplied(h, w, 7)
    for rule in (0, 1):
        for col in (pack(255, 255, 255, 255), pack(60, 50, 40, 120)):
            b = FillBatch()
            b.add(segs, col, rule, MaskBlend)
            b.add(segs, pack(10, 200, 30, 255), rule, NormalBlend)
            _check(b.arrays(), w, h, background=bg)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(512, 512, 3), (333, 97, 2), (4100, 40, 1)])
def test_run_cleared_equals_clear_then_run(size):
    """pixie_cuda_cmdlist_run_cleared (the raster kernel zeroes each row tile before its first fill) on a DIRTY canvas
    == image.fill(0) + run, incl. widths that are not a multiple of 4, two column tiles, several layers, MaskBlend
    fills (which clear what they do not cover) and rows / layers no fill touches."""
    from pixie_b200 import device as dev

    w, h, layers = size
    dev.init(0)
    b = FillBatch()
    for layer in range(layers):  # fills in layer order
        if layer == 1:
            continue  # a layer without fills: must still come out transparent
        synth.icon_fills(7 + layer, min(w, h) if min(w, h) >= 64 else 64, layer, b)
        if layer == 0:
            b.add(host.fill_segments(f"M 3.5 2.5 L {w - 7.25} {h * 0.4} L {w * 0.3} {h - 3.5} z"), pack(10, 200, 30, 200), host.NonZero, NormalBlend, 0)
    b.add(host.fill_segments(f"M -20 {h * 0.2} H {w * 0.6} V {h * 0.7} H -20 z"), pack(0, 0, 0, 180), host.NonZero, MaskBlend, layers - 1)
    arrays = b.arrays()
    dirty = np.stack([synth.random_premultiplied(h, w, 40 + k) for k in range(layers)]) if layers > 1 else synth.random_premultiplied(h, w, 40)
    a = dev.DeviceImage(w, h, layers).upload(dirty)
    ref = dev.DeviceImage(w, h, layers).upload(dirty)
    cl = dev.CmdList(w, h, layers, arrays)
    ca = cl.run(a, count_covered=True, clear=True)
    ref.fill(0)
    cb = cl.run(ref, count_covered=True)
    assert ca == cb
    got, want = a.download(), ref.download()
    assert np.array_equal(got, want)
    ow, _ = oracle_render_batch(arrays, w, h, layers=layers)
    assert diff_report(got.reshape(layers, h, w, 4)[0], ow[0])[0] == 0
