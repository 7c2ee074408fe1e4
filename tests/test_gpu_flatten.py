"""Device-side flattening + stroking (flatten.cu) against libpixie_host.so, whose
commandsToShapes / strokeShapes / shapesToSegments are pinned on the reference's goldens: the segment arrays a command
list holds after pixie_cuda_cmdlist_create_from_paths must EQUAL the host's (same floats, same order, same windings),
and so must the rendered pixels."""
import os

import numpy as np
import pytest

from pixie_b200 import host, svg as psvg, synth
from pixie_b200.device import FillBatch, PathBatch

pytestmark = pytest.mark.gpu

TIGER = os.path.join(os.path.dirname(__file__), "golden", "tiger.svg")

PATHS = [
    "M 10 10 L 90 10 L 90 90 L 10 90 Z",
    "M 20.5 30.25 l 40 3 l -7 50.5 z m 5 5 h 10 v 10 h -10 z",
    "M 10 80 C 40 10, 65 10, 95 80 S 150 150, 180 80",
    "M 10 80 Q 52.5 10, 95 80 T 180 80 T 250 20",
    "m 30 30 c 10 -20 30 -20 40 0 s 30 20 40 0 q 10 -30 20 0 t 20 0 z",
    "M 0.5 0.5 L 100.25 3.125 L 50 75.75 Z M 200 200 L 220 200 L 220 220",
    "M 5 5 L 5 5 L 60 5 L 60 5 L 60 60 Z",                      # zero-length segments
    "M 100 100 C 100 100 100 100 100 100 L 150 160 Q 150 160 150 160 Z",  # degenerate curves
    "M 10 10 L 50 50 Z L 90 10 L 90 50 Z",                        # drawing continues after Close without a Move
    "M 300 10 H 350 V 60 H 300 Z M 310 20 V 50 H 340 V 20 Z",
    "M 3 200 C 3 500 400 -200 401 300 C 700 900 -300 100 20 20",  # wild cubic: many halvings
]


def _transforms():
    f = np.float32
    return [None, host.scale(f(2.5), f(0.75)), host.matmul(host.translate(f(13.25), f(-7.5)), host.rotate(f(0.37))),
            host.matmul(host.scale(f(0.04), f(0.04)), host.translate(f(500), f(500)))]


def _host_batch(items):
    b = FillBatch()
    for it in items:
        if it[0] == "fill":
            _, path, tr, rgbx, rule, mode = it
            b.add(host.fill_segments(path, tr), rgbx, rule, mode, 0)
        else:
            _, path, tr, sw, cap, join, miter, rgbx = it
            b.add(host.stroke_segments(path, tr, sw, cap, join, miter, ()), rgbx, 0, 0, 0)
    return b.arrays()


def _path_batch(items):
    b = PathBatch()
    for it in items:
        if it[0] == "fill":
            _, path, tr, rgbx, rule, mode = it
            b.add_fill(path, tr, rgbx, rule, mode, 0)
        else:
            _, path, tr, sw, cap, join, miter, rgbx = it
            b.add_stroke(path, tr, sw, cap, join, miter, (), rgbx, 0, 0, 0)
    return b


def _assert_same_segments(cl, want):
    xy, wd, so = cl.segments()
    assert so.tolist() == want["seg_offsets"].tolist()
    assert wd.tolist() == want["winding"].tolist()
    # float equality (a -0.0 where the host has +0.0 is the same coordinate)
    assert xy.shape == want["xyxy"].shape
    bad = np.argwhere(~(xy == want["xyxy"]))
    assert len(bad) == 0, (bad[:5], xy[bad[0][0]], want["xyxy"][bad[0][0]])


def _items():
    items = []
    col = 0xFF2040C0
    for k, d in enumerate(PATHS):
        for t, tr in enumerate(_transforms()):
            p = host.parsePath(d)
            items.append(("fill", p, tr, col + 0x010101 * (k + t), (k + t) & 1, 0))
            for cap in (host.ButtCap, host.SquareCap):
                for join, miter in ((host.MiterJoin, 4.0), (host.MiterJoin, 1.2), (host.BevelJoin, 4.0)):
                    items.append(("stroke", p, tr, 3.5 + k * 0.75, cap, join, miter, 0xC0102030 + k))
    return items


def test_segments_equal_host_flattener():
    from pixie_b200 import device as dev

    dev.init(0)
    items = _items()
    want = _host_batch(items)
    cl = dev.CmdList.from_paths(512, 512, 1, _path_batch(items))
    assert cl.info()["segments"] == len(want["winding"])
    _assert_same_segments(cl, want)


def test_mixed_with_host_fallback_paths():
    """Arcs, round joins / caps and dashes travel as finished segments between device-flattened paths."""
    from pixie_b200 import device as dev

    dev.init(0)
    circle = host.newPath()
    circle.circle(100, 100, 60)
    tri = host.parsePath("M 20 20 L 200 40 L 90 180 Z")
    arcp = host.parsePath("M 10 100 A 50 30 20 1 0 200 120 L 100 10 Z")
    b, hb = PathBatch(), FillBatch()
    b.add_fill(tri, None, 0xFF0000FF, 0, 17, 0)
    hb.add(host.fill_segments(tri, None), 0xFF0000FF, 0, 17, 0)
    b.add_fill(circle, None, 0x80004000, 0, 0, 0)
    hb.add(host.fill_segments(circle, None), 0x80004000, 0, 0, 0)
    b.add_stroke(tri, None, 6.0, host.RoundCap, host.RoundJoin, 4.0, (), 0xFF00FF00, 0, 0, 0)
    hb.add(host.stroke_segments(tri, None, 6.0, host.RoundCap, host.RoundJoin, 4.0, ()), 0xFF00FF00, 0, 0, 0)
    b.add_stroke(tri, None, 4.0, host.ButtCap, host.MiterJoin, 4.0, (5.0, 3.0), 0xFFFF0000, 0, 0, 0)
    hb.add(host.stroke_segments(tri, None, 4.0, host.ButtCap, host.MiterJoin, 4.0, (5.0, 3.0)), 0xFFFF0000, 0, 0, 0)
    b.add_fill(arcp, None, 0xFF808080, 1, 0, 0)
    hb.add(host.fill_segments(arcp, None), 0xFF808080, 1, 0, 0)
    b.add_stroke(tri, host.scale(np.float32(1.5), np.float32(1.5)), 3.0, host.SquareCap, host.BevelJoin, 4.0, (), 0xFF123456, 0, 0, 0)
    hb.add(host.stroke_segments(tri, host.scale(np.float32(1.5), np.float32(1.5)), 3.0, host.SquareCap, host.BevelJoin, 4.0, ()),
           0xFF123456, 0, 0, 0)
    assert b.host_paths == 3  # round joins, dashes, the arc path (a circle is four cubics)
    want = hb.arrays()
    cl = dev.CmdList.from_paths(256, 256, 1, b)
    _assert_same_segments(cl, want)
    a, c = dev.DeviceImage(256, 256), dev.DeviceImage(256, 256)
    cl.run(a)
    dev.CmdList(256, 256, 1, want).run(c)
    assert (a.download() == c.download()).all()


@pytest.mark.parametrize("size", [900, 4096])
def test_tiger_from_commands(size):
    """The whole tiger from path commands: 305 fills / strokes flattened on the device, pixels equal to the host-flattened
    render (which the other tiger tests pin against the oracle)."""
    from pixie_b200 import device as dev

    dev.init(0)
    doc = psvg.parseSvg(open(TIGER).read(), size, size)
    want = psvg.svg_fill_batch(doc).arrays()
    pb = psvg.svg_path_batch(doc)
    cl = dev.CmdList.from_paths(size, size, 1, pb)
    _assert_same_segments(cl, want)
    a, c = dev.DeviceImage(size, size), dev.DeviceImage(size, size)
    cov = cl.run(a, count_covered=True)
    cov2 = dev.CmdList(size, size, 1, want).run(c, count_covered=True)
    assert cov == cov2
    assert a.checksum() == c.checksum()


def test_synthetic_icons_from_commands():
    """Icon-like documents on separate layers (BASELINE config 5 in miniature)."""
    from pixie_b200 import device as dev

    dev.init(0)
    rng = np.random.default_rng(7)
    b, hb = PathBatch(), FillBatch()
    n = 24
    for layer in range(n):
        for s in range(4):
            p = host.newPath()
            x, y = rng.uniform(8, 56, 2)
            p.moveTo(x, y)
            for _ in range(int(rng.integers(2, 7))):
                kind = int(rng.integers(0, 3))
                q = rng.uniform(2, 62, 6).astype(np.float32)
                if kind == 0:
                    p.lineTo(q[0], q[1])
                elif kind == 1:
                    p.quadraticCurveTo(q[0], q[1], q[2], q[3])
                else:
                    p.bezierCurveTo(*q)
            if s & 1:
                p.closePath()
            col = int(rng.integers(0, 1 << 32))
            a8 = col >> 24
            col = (col & 0xFF000000) | (((col & 255) * a8 // 255)) | ((((col >> 8) & 255) * a8 // 255) << 8) | ((((col >> 16) & 255) * a8 // 255) << 16)
            if s < 2:
                b.add_fill(p, None, col, s & 1, 0, layer)
                hb.add(host.fill_segments(p, None), col, s & 1, 0, layer)
            else:
                b.add_stroke(p, None, 2.5, host.SquareCap if s == 2 else host.ButtCap, host.MiterJoin, 4.0, (), col, 0, 0, layer)
                hb.add(host.stroke_segments(p, None, 2.5, host.SquareCap if s == 2 else host.ButtCap, host.MiterJoin, 4.0, ()), col, 0, 0, layer)
    want = hb.arrays()
    cl = dev.CmdList.from_paths(64, 64, n, b)
    _assert_same_segments(cl, want)
    a, c = dev.DeviceImage(64, 64, n), dev.DeviceImage(64, 64, n)
    cl.run(a)
    dev.CmdList(64, 64, n, want).run(c)
    assert (a.download() == c.download()).all()


def test_errors():
    from pixie_b200 import device as dev
    from pixie_b200.common import PixieError

    dev.init(0)
    b = PathBatch()
    b.add_fill(host.parsePath("M 0 0 L 10 0 L 10 10 Z"), None, 0xFF0000FF, 0, 0, 0)
    b.descs[0].line_join = 1
    b.descs[0].kind = 1
    with pytest.raises(PixieError):
        dev.CmdList.from_paths(32, 32, 1, b)
    b = PathBatch()
    b.add_fill(host.parsePath("M 0 0 L 10 0 L 10 10 Z"), None, 0xFF0000FF, 0, 0, 0)
    b.descs[0].num_commands = 1  # fewer slots than the stream needs
    with pytest.raises(PixieError):
        dev.CmdList.from_paths(32, 32, 1, b)
    empty = dev.CmdList.from_paths(32, 32, 1, PathBatch())
    img = dev.DeviceImage(32, 32)
    assert empty.run(img, count_covered=True) == 0


def test_render_paths_host_and_read_svg():
    """pixie_cuda_render_paths_host (commands in, host pixels out, banded D2H) and api.readSvg give the pixels of the
    host-flattened render."""
    from pixie_b200 import api, device as dev

    size = 2048
    dev.init(0)
    data = open(TIGER).read()
    doc = psvg.parseSvg(data, size, size)
    want_img = dev.DeviceImage(size, size)
    cov_want = dev.CmdList(size, size, 1, psvg.svg_fill_batch(doc).arrays()).run(want_img, count_covered=True)
    want = want_img.download()
    pb = psvg.svg_path_batch(doc)
    pinned = dev.PinnedBuffer(size * size * 4)
    cov = dev.render_paths_host(pinned.ptr, size, size, pb, count_covered=True)
    assert cov == cov_want
    assert (pinned.array[:size * size * 4].reshape(size, size, 4) == want).all()
    pageable = np.full((size, size, 4), 7, np.uint8)
    dev.render_paths_host(pageable.ctypes.data, size, size, pb)
    assert (pageable == want).all()
    # drawing over existing pixels (clear = False): NormalBlend paths over an opaque backdrop
    b = PathBatch()
    tri = host.parsePath("M 100 100 L 1900 300 L 700 1800 Z")
    b.add_fill(tri, None, 0x80402010, 0, 0, 0)
    b.add_stroke(tri, None, 25.0, host.SquareCap, host.BevelJoin, 4.0, (), 0xFF00FF00, 0, 0, 0)
    back = np.full((size, size, 4), 200, np.uint8)
    back[..., 3] = 255
    got = back.copy()
    dev.render_paths_host(got.ctypes.data, size, size, b, clear=False)
    ref = dev.DeviceImage(size, size).upload(back)
    hb = FillBatch()
    hb.add(host.fill_segments(tri, None), 0x80402010, 0, 0, 0)
    hb.add(host.stroke_segments(tri, None, 25.0, host.SquareCap, host.BevelJoin, 4.0, ()), 0xFF00FF00, 0, 0, 0)
    dev.fill_batch(ref, hb.arrays())
    assert (got == ref.download()).all()
    img = api.readSvg(data, size, size)
    assert (img._d.download() == want).all()
