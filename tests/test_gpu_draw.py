"""GPU parity of draw-with-transform, minify / magnify and the gradient paints (through the C ABI) against
the CPU oracle on seeded inputs: bit-exact for every blend mode, transform class and odd size.
(images.nim:168-259, :367-449, :531-683; paints.nim:68-248)"""
import math

import numpy as np
import pytest

from pixie_b200 import host, synth
from pixie_b200.common import MaskBlend, NormalBlend, OverwriteBlend

pytestmark = pytest.mark.gpu

f = np.float32


def _backends():
    from _gpu_backend import GpuBackend
    from _oracle import OracleBackend

    return GpuBackend(), OracleBackend(0)


def _rand(h, w, seed):
    return synth.random_premultiplied(h, w, seed)


def _T(x, y):
    return host.translate(f(x), f(y))


def _mats():
    r = lambda deg: host.rotate(f(f(deg) * f(math.pi) / f(180)))
    return {
        "identity": host.mat3(),
        "int_translate": _T(7, -3),
        "frac_translate": _T(10.25, 3.75),
        "rotate30": host.matmul(_T(40, 10), r(30)),
        "rotate-75_scale": host.matmul(host.matmul(_T(5, 60), r(-75)), host.scale(f(1.3), f(0.8))),
        "scale_half": host.matmul(_T(3, 3), host.scale(f(0.5), f(0.5))),
        "scale_0.2": host.scale(f(0.2), f(0.2)),
        "scale_3.7": host.matmul(_T(-20, -30), host.scale(f(3.7), f(3.7))),
        "scale_0.07_rot": host.matmul(host.matmul(_T(50, 50), r(12)), host.scale(f(0.07), f(0.07))),
        "shear": np.array([1, 0.3, 0, -0.2, 1, 0, 12, 5, 1], np.float32),
        "flip": host.matmul(_T(90, 0), host.scale(f(-1), f(1))),
        "offscreen": _T(500, 500),
    }


@pytest.mark.parametrize("name", sorted(_mats()))
@pytest.mark.parametrize("mode", [NormalBlend, OverwriteBlend, MaskBlend, 2, 8, 11, 19])
def test_draw_transforms(name, mode):
    gb, ob = _backends()
    src = _rand(97, 131, 5)
    base = _rand(101, 127, 6)
    a, b = base.copy(), base.copy()
    gb.draw(a, src, _mats()[name], mode)
    ob.draw(b, src, _mats()[name], mode)
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() == 0, f"{name} mode {mode}: {(d.max(-1) > 0).sum()} px differ, max {d.max()}"


@pytest.mark.parametrize("mode", list(range(20)))
def test_draw_smooth_all_modes(mode):
    gb, ob = _backends()
    src = _rand(64, 64, 7 + mode)
    base = _rand(80, 96, 8)
    m = host.matmul(_T(20.5, 9.25), host.rotate(f(0.4)))
    a, b = base.copy(), base.copy()
    gb.draw(a, src, m, mode)
    ob.draw(b, src, m, mode)
    assert np.array_equal(a, b), f"mode {mode}: {(np.abs(a.astype(int) - b.astype(int)).max(-1) > 0).sum()} px differ"


def test_draw_large_wide_rows():
    """Rows longer than one warp pass many times: the position accumulates over thousands of additions."""
    gb, ob = _backends()
    src = _rand(300, 2100, 9)
    base = _rand(256, 4100, 10)
    m = host.matmul(_T(3.3, -20.7), host.scale(f(1.9), f(1.1)))
    a, b = base.copy(), base.copy()
    gb.draw(a, src, m, NormalBlend)
    ob.draw(b, src, m, NormalBlend)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("tiled", [True, False])
@pytest.mark.parametrize("mode", [NormalBlend, 5, MaskBlend])
def test_draw_correct(tiled, mode):
    gb, ob = _backends()
    src = _rand(37, 53, 11)
    base = _rand(90, 110, 12)
    for m in (host.scale(f(0.3), f(0.3)), host.matmul(_T(13.5, 4.5), host.rotate(f(-0.3))), host.scale(f(2.5), f(2.5)),
              _T(-30.5, -20.25)):
        a, b = base.copy(), base.copy()
        gb.draw_tiled(a, src, m, mode, tiled)
        ob.draw_tiled(b, src, m, mode, tiled)
        assert np.array_equal(a, b), f"tiled={tiled} mode={mode}"


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 5), (99, 101), (100, 100), (257, 64), (64, 257), (1, 9), (9, 1)])
def test_minify_magnify(shape):
    gb, ob = _backends()
    img = _rand(shape[0], shape[1], 13)
    for power in (0, 1, 2, 3):
        assert np.array_equal(gb.minify_by2(img, power), ob.minify_by2(img, power)), f"minify {shape} power {power}"
    for power in (0, 1, 2):
        assert np.array_equal(gb.magnify_by2(img, power), ob.magnify_by2(img, power)), f"magnify {shape} power {power}"


def test_minify_negative_power_raises():
    from pixie_b200 import device as dev
    from pixie_b200.common import PixieError

    dev.init(0)
    d = dev.DeviceImage(4, 4)
    with pytest.raises(PixieError, match="Cannot minifyBy2 with negative power"):
        dev.minify_by2(d, -1)
    with pytest.raises(PixieError, match="Cannot magnifyBy2 with negative power"):
        dev.magnify_by2(d, -1)


_STOPS3 = [(0.0, (1, 0, 0, 1)), (0.3, (0, 1, 0, 0.5)), (0.8, (0.2, 0.4, 1, 1)), (1.0, (1, 1, 1, 0.1))]


@pytest.mark.parametrize("kind,handles", [
    (3, [(0, 50), (100, 50)]), (3, [(50, 0), (50, 100)]), (3, [(10, 20), (150, 90)]), (3, [(120, 80), (-10, 5)]),
    (4, [(50, 50), (100, 50), (50, 100)]), (4, [(70, 40), (120, 90), (30, 60)]),
    (5, [(50, 50), (100, 50), (50, 100)]), (5, [(81.5, 33.25), (10, 90), (0, 0)]),
])
@pytest.mark.parametrize("opacity", [1.0, 0.37])
def test_gradients(kind, handles, opacity):
    gb, ob = _backends()
    for shape in ((100, 100), (77, 203)):
        a = _rand(shape[0], shape[1], 14)
        b = a.copy()
        gb.fill_gradient(a, kind, handles, _STOPS3, opacity)
        ob.fill_gradient(b, kind, handles, _STOPS3, opacity)
        d = np.abs(a.astype(int) - b.astype(int))
        assert d.max() == 0, f"kind {kind} {handles}: {(d.max(-1) > 0).sum()} px differ, max {d.max()}"


def test_gradient_errors_and_single_stop():
    from pixie_b200 import device as dev
    from pixie_b200.common import PixieError

    gb, ob = _backends()
    d = dev.DeviceImage(8, 8)
    with pytest.raises(PixieError, match="Linear gradient requires 2 handles"):
        dev.fill_gradient(d, 3, [(0, 0)], _STOPS3)
    with pytest.raises(PixieError, match="Radial gradient requires 3 handles"):
        dev.fill_gradient(d, 4, [(0, 0), (1, 1)], _STOPS3)
    with pytest.raises(PixieError, match="Gradient must have at least 1 color stop"):
        dev.fill_gradient(d, 3, [(0, 0), (1, 1)], [])
    a = np.zeros((16, 16, 4), np.uint8)
    b = a.copy()
    gb.fill_gradient(a, 3, [(0, 0), (16, 16)], [(0.5, (0.2, 0.5, 0.7, 0.6))])
    ob.fill_gradient(b, 3, [(0, 0), (16, 16)], [(0.5, (0.2, 0.5, 0.7, 0.6))])
    assert np.array_equal(a, b) and a.any()


def test_shadow_fractional_offset():
    """shadow() with a fractional offset: the offset copy goes through drawSmooth (images.nim:768-769)."""
    gb, ob = _backends()
    img = np.zeros((90, 120, 4), np.uint8)
    img[20:60, 30:90] = (255, 255, 255, 255)
    lut = host.gaussianKernel(6)
    got = gb.shadow(img, 3.5, -2.25, 2, lut, 6, 0xC8000000)
    want = ob.shadow(img, 3.5, -2.25, 2, lut, 6, 0xC8000000)
    assert np.array_equal(got, want)


def test_api_paints_match_reference_goldens():
    """tests/test_paints.nim written against pixie_b200.api: gradient and image paints, byte for byte."""
    import golden_cases as gc
    from pixie_b200 import api as pixie

    stops = [pixie.ColorStop((1, 0, 0, 1), 0.0), pixie.ColorStop((1, 0, 0, 0.15625), 1.0)]
    for golden, kind, handles, opacity in [
        ("paths_gradientLinear.png", pixie.LinearGradientPaint, [(0, 50), (100, 50)], 1.0),
        ("paths_gradientLinear2.png", pixie.LinearGradientPaint, [(50, 0), (50, 100)], 1.0),
        ("paths_gradientRadial.png", pixie.RadialGradientPaint, [(50, 50), (100, 50), (50, 100)], 1.0),
        ("paths_gradientAngular.png", pixie.AngularGradientPaint, [(50, 50), (100, 50), (50, 100)], 1.0),
        ("paths_gradientAngularOpacity.png", pixie.AngularGradientPaint, [(50, 50), (100, 50), (50, 100)], 0.5),
    ]:
        paint = pixie.newPaint(kind)
        paint.gradientHandlePositions = handles
        paint.gradientStops = stops
        paint.opacity = opacity
        image = pixie.newImage(100, 100)
        image.fillPath(gc.HEART_PAINT, paint)
        assert gc.compare(image.data, gc.load_golden(golden)) == (0, 0), golden

    mandrill = pixie.newImage(512, 512)
    mandrill.data = gc.load_golden("fileformats_png_mandrill.png")
    for golden, kind, s, opacity in [
        ("paths_paintImage.png", pixie.ImagePaint, 0.2, 1.0), ("paths_paintImageOpacity.png", pixie.ImagePaint, 0.2, 0.5),
        ("paths_paintImageTiled.png", pixie.TiledImagePaint, 0.02, 1.0),
        ("paths_paintImageTiledOpacity.png", pixie.TiledImagePaint, 0.02, 0.5),
    ]:
        paint = pixie.newPaint(kind)
        paint.image = mandrill
        paint.imageMat = host.scale(f(s), f(s))
        paint.opacity = opacity
        image = pixie.newImage(100, 100)
        image.fillPath(gc.HEART_PAINT, paint)
        assert gc.compare(image.data, gc.load_golden(golden)) == (0, 0), golden

    for golden, kind, s in [("paths_fillImagePaint.png", pixie.ImagePaint, 0.2),
                            ("paths_fillTiledImagePaint.png", pixie.TiledImagePaint, 0.1)]:
        paint = pixie.newPaint(kind)
        paint.image = mandrill
        paint.imageMat = host.scale(f(s), f(s))
        paint.opacity = 0.5
        image = pixie.newImage(128, 128)
        image.fill((0, 255, 0, 255))
        image.fill(paint)
        assert gc.compare(image.data, gc.load_golden(golden)) == (0, 0), golden


def test_api_draw_rotate_and_resize():
    import golden_cases as gc
    from pixie_b200 import api as pixie

    a = pixie.newImage(1000, 1000)
    b = pixie.newImage(500, 500)
    a.fill((255, 0, 0, 255))
    b.fill((0, 255, 0, 255))
    a.draw(b, host.matmul(_T(250, 250), host.rotate(f(f(-90) * f(math.pi) / f(180)))))
    assert gc.compare(a.data, gc.load_golden("images_rotate90.png")) == (0, 0)
    small = b.resize(100, 60)
    assert (small.width, small.height) == (100, 60)
    assert tuple(small[50, 30]) == (0, 255, 0, 255)
    mini = a.minifyBy2(2)
    assert (mini.width, mini.height) == (250, 250)


def test_draw_random_transforms_fuzz():
    """The closed-form position chain (draw.cu build_chain) against the oracle's sequential `srcPos += dx` on
    random affine transforms, including near-axis-aligned rotations, dyadic scales (rounding ties) and rows that
    cross zero."""
    gb, ob = _backends()
    rng = np.random.default_rng(2026)
    src = _rand(120, 150, 21)
    base = _rand(140, 1300, 22)
    for trial in range(40):
        kind = trial % 5
        if kind == 0:
            m = host.matmul(_T(rng.uniform(-200, 900), rng.uniform(-100, 100)), host.rotate(f(rng.uniform(-3.2, 3.2))))
        elif kind == 1:
            m = host.matmul(_T(rng.uniform(0, 600), rng.uniform(0, 60)), host.rotate(f(rng.uniform(-1e-3, 1e-3))))
        elif kind == 2:
            s_ = f(2.0 ** rng.integers(-1, 2)) * f(rng.choice([0.75, 1.0, 1.25, 1.5]))
            m = host.matmul(_T(rng.integers(-50, 700) + rng.choice([0.0, 0.5, 0.25]), rng.integers(-20, 40)), host.scale(s_, s_))
        elif kind == 3:
            m = np.array([rng.uniform(0.6, 1.9), rng.uniform(-0.5, 0.5), 0, rng.uniform(-0.5, 0.5), rng.uniform(0.6, 1.9), 0,
                          rng.uniform(-100, 800), rng.uniform(-60, 60), 1], np.float32)
        else:
            m = host.matmul(host.matmul(_T(rng.uniform(300, 900), rng.uniform(20, 100)), host.rotate(f(rng.uniform(-3.2, 3.2)))),
                            host.scale(f(rng.uniform(0.55, 1.9)), f(rng.uniform(0.55, 1.9))))
        a, b = base.copy(), base.copy()
        gb.draw(a, src, m, NormalBlend)
        ob.draw(b, src, m, NormalBlend)
        assert np.array_equal(a, b), f"trial {trial} kind {kind}: {(np.abs(a.astype(int) - b.astype(int)).max(-1) > 0).sum()} px differ"


def test_host_variants_equal_handle_path():
    """pixie_cuda_draw_host / fill_gradient_host / minify_by2_host / magnify_by2_host (what the Nim shim binds)."""
    from pixie_b200 import device as dev

    gb, ob = _backends()
    L = dev.lib()
    src = _rand(60, 70, 31)
    base = _rand(90, 110, 32)
    m = np.ascontiguousarray(host.matmul(_T(12.5, 7.25), host.rotate(f(0.3))), np.float32)
    for tiled in (0, 1):
        got, want = base.copy(), base.copy()
        dev.check(L.pixie_cuda_draw_host(got.ctypes.data, 110, 90, src.ctypes.data, 70, 60, m.ctypes.data, NormalBlend, tiled))
        (ob.draw_tiled if tiled else ob.draw)(want, src, m, NormalBlend)
        assert np.array_equal(got, want)
    got, want = base.copy(), base.copy()
    hx = np.array([10, 20, 90, 70], np.float32)
    pos = np.array([0.0, 1.0], np.float32)
    col = np.array([1, 0, 0, 1, 0, 0, 1, 0.25], np.float32)
    dev.check(L.pixie_cuda_fill_gradient_host(got.ctypes.data, 110, 90, 3, hx.ctypes.data, 2, pos.ctypes.data, col.ctypes.data, 2, 0.8))
    ob.fill_gradient(want, 3, [(10, 20), (90, 70)], [(0.0, (1, 0, 0, 1)), (1.0, (0, 0, 1, 0.25))], 0.8)
    assert np.array_equal(got, want)
    mini = np.zeros((45, 55, 4), np.uint8)
    dev.check(L.pixie_cuda_minify_by2_host(base.ctypes.data, 110, 90, 1, mini.ctypes.data))
    assert np.array_equal(mini, ob.minify_by2(base, 1))
    big = np.zeros((180, 220, 4), np.uint8)
    dev.check(L.pixie_cuda_magnify_by2_host(base.ctypes.data, 110, 90, 1, big.ctypes.data))
    assert np.array_equal(big, ob.magnify_by2(base, 1))


@pytest.mark.parametrize("mat", [
    [0, 0, 0, 0, 0, 0, 5, 5, 1],          # singular: the inverse is inf / NaN everywhere
    [1, 0, 0, 0, 0, 0, 3, 4, 1],          # rank 1
    [1e-30, 0, 0, 0, 1e-30, 0, 10, 10, 1],  # astronomically small scale: magnify loop is bounded by the size check
    [1e6, 0, 0, 0, 1e6, 0, -3e7, -2e7, 1],  # one source pixel covers everything
], ids=["zero", "rank1", "tiny", "huge"])
def test_draw_degenerate_transforms(mat):
    """Degenerate transforms must neither hang nor crash, and agree with the oracle where it has an answer."""
    from pixie_b200.common import PixieError

    gb, ob = _backends()
    src = _rand(8, 9, 41)
    base = _rand(40, 50, 42)
    m = np.array(mat, np.float32)
    a, b = base.copy(), base.copy()
    from _oracle import OracleError

    try:
        ob.draw(b, src, m, NormalBlend)
    except OracleError:  # the source would have to be magnified beyond 2^28 pixels (the reference runs out of memory)
        with pytest.raises(PixieError, match="too large"):
            gb.draw(a, src, m, NormalBlend)
        return
    gb.draw(a, src, m, NormalBlend)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("kind", [3, 4, 5])
@pytest.mark.parametrize("shape", [(256, 96), (203, 61)])
def test_fill_gradient_masked_equals_unfused_sequence(kind, shape):
    """pixie_cuda_fill_gradient_masked == fillGradient(opacity 1) + applyOpacity(mask) + blend_rect_masked, for every
    blend mode, RGBX and A8 masks, paint opacities 1 / 0.6 (paths.nim:2115-2142)."""
    from pixie_b200 import device as dev, synth

    w, h = shape
    dev.init(0)
    rng = np.random.default_rng(kind * 100 + w)
    dst0 = synth.random_premultiplied(h, w, 5)
    cov = synth.coverage_mask(h, w, 6)
    cov[:, : w // 5] = 0       # uncovered stretches (the skip path) and fully covered ones
    cov[: h // 4, w // 2:] = 255
    mask_rgbx = np.zeros((h, w, 4), np.uint8)
    mask_rgbx[..., :] = cov[..., None]
    handles = [(w * 0.2, h * 0.3), (w * 0.9, h * 0.7)] if kind == 3 else [(w * 0.5, h * 0.5), (w * 0.9, h * 0.5), (w * 0.5, h * 0.95)]
    stops = [(0.0, (1.0, 0.2, 0.1, 1.0)), (0.35, (0.1, 0.9, 0.3, 0.4)), (1.0, (0.2, 0.1, 1.0, 0.85))]
    for mode in range(20):
        for opacity in (1.0, 0.6):
            for a8 in (False, True):
                want = dev.DeviceImage(w, h).upload(dst0)
                fill = dev.DeviceImage(w, h)
                dev.fill_gradient(fill, kind, handles, stops, 1.0)
                m = dev.DeviceImage(w, h).upload(mask_rgbx)
                if opacity != 1.0:
                    dev.apply_opacity(m, opacity)
                dev.blend_rect_masked(want, fill, m, 0, 0, mode)
                got = dev.DeviceImage(w, h).upload(dst0)
                if a8:
                    m2 = dev.DeviceImage(w, h, a8=True).upload(np.ascontiguousarray(cov))
                else:
                    m2 = dev.DeviceImage(w, h).upload(mask_rgbx)
                dev.fill_gradient_masked(got, m2, kind, handles, stops, opacity, mode)
                a, b = got.download(), want.download()
                assert (a == b).all(), (mode, opacity, a8, int((a != b).any(axis=-1).sum()))
