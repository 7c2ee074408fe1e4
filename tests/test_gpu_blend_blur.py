"""Parity of blend_rect (all 20 modes, masks, clipping), blur, spread and shadow with the oracle."""
import numpy as np
import pytest

from pixie_b200 import host, synth
from pixie_b200.common import BLEND_MODE_NAMES, MaskBlend, NormalBlend, rgbx as pack
from _util import diff_report

pytestmark = pytest.mark.gpu


def _backends():
    from _gpu_backend import GpuBackend
    from _oracle import OracleBackend

    return GpuBackend(), OracleBackend(0)


@pytest.mark.parametrize("mode", range(20), ids=BLEND_MODE_NAMES)
def test_blend_rect_all_modes(mode):
    gb, ob = _backends()
    for (dw, dh, sw, sh, px, py) in [(256, 64, 256, 64, 0, 0), (200, 50, 120, 40, 36, 7), (131, 37, 64, 64, -13, -9),
                                     (96, 96, 50, 50, 70, 80), (64, 16, 64, 16, 4, 0), (33, 9, 200, 100, -100, -50)]:
        dst = synth.random_premultiplied(dh, dw, 11 + mode)
        src = synth.random_premultiplied(sh, sw, 77 + mode)
        a, b = dst.copy(), dst.copy()
        gb.blend_rect(a, src, px, py, mode)
        ob.blend_rect(b, src, px, py, mode)
        n, mx, where = diff_report(a, b)
        assert n == 0, f"{BLEND_MODE_NAMES[mode]} {(dw, dh, sw, sh, px, py)}: {n} px differ (max {mx}) at {where}"


@pytest.mark.parametrize("mode", range(20), ids=BLEND_MODE_NAMES)
def test_blend_rect_non_premultiplied_inputs(mode):
    """Arbitrary bytes (not valid premultiplied colours): uint8 wrap-around is part of the semantics."""
    gb, ob = _backends()
    rng = np.random.default_rng(mode)
    dst = rng.integers(0, 256, (32, 256, 4), dtype=np.uint8)
    src = rng.integers(0, 256, (32, 256, 4), dtype=np.uint8)
    a, b = dst.copy(), dst.copy()
    gb.blend_rect(a, src, 0, 0, mode)
    ob.blend_rect(b, src, 0, 0, mode)
    n, mx, where = diff_report(a, b)
    assert n == 0, f"{BLEND_MODE_NAMES[mode]}: {n} px differ (max {mx}) at {where}"


@pytest.mark.parametrize("mode", [0, 2, 3, 8, 11, 16, 17, 19])
@pytest.mark.parametrize("rgbx_mask", [False, True])
def test_blend_rect_masked(mode, rgbx_mask):
    """Fused mask*fill composite (paths.nim:2141-2142) == MaskBlend draw followed by the blend draw."""
    gb, ob = _backends()
    for (dw, dh, sw, sh, px, py) in [(256, 32, 256, 32, 0, 0), (100, 40, 64, 30, 19, 5), (64, 64, 64, 64, -8, 8)]:
        dst = synth.random_premultiplied(dh, dw, 1)
        src = synth.random_premultiplied(sh, sw, 2)
        cov = synth.coverage_mask(sh, sw, 3)
        mask = cov
        if rgbx_mask:
            mask = np.zeros((sh, sw, 4), np.uint8)
            mask[..., 3] = cov
            mask[..., :3] = cov[..., None] // 2
        a, b = dst.copy(), dst.copy()
        gb.blend_rect_masked(a, src, mask, px, py, mode)
        ob.blend_rect_masked(b, src, mask, px, py, mode)
        assert diff_report(a, b)[0] == 0
        # and the oracle's fused form equals the reference's two draws
        tmp = src.copy()
        m4 = mask if rgbx_mask else np.concatenate([np.zeros((sh, sw, 3), np.uint8), cov[..., None]], -1)
        ob.blend_rect(tmp, np.ascontiguousarray(m4), 0, 0, MaskBlend)
        c = dst.copy()
        ob.blend_rect(c, tmp, px, py, mode)
        assert diff_report(b, c)[0] == 0


def test_blend_px_exhaustive_sample():
    """Every mode on a dense lattice of (backdrop, source) channel/alpha values via 1-row images."""
    gb, ob = _backends()
    vals = np.array([0, 1, 2, 63, 64, 127, 128, 129, 191, 254, 255], np.uint8)
    a_ = np.array([0, 1, 127, 128, 254, 255], np.uint8)
    bc, ba, sc, sa = np.meshgrid(vals, a_, vals, a_, indexing="ij")
    n = bc.size
    dst = np.stack([bc.ravel(), 255 - bc.ravel(), bc.ravel() // 2, ba.ravel()], -1).reshape(1, n, 4).astype(np.uint8)
    src = np.stack([sc.ravel(), sc.ravel() // 3, 255 - sc.ravel(), sa.ravel()], -1).reshape(1, n, 4).astype(np.uint8)
    pad = (-n) % 4
    dst = np.ascontiguousarray(np.pad(dst, ((0, 0), (0, pad), (0, 0))))
    src = np.ascontiguousarray(np.pad(src, ((0, 0), (0, pad), (0, 0))))
    for mode in range(20):
        a, b = dst.copy(), dst.copy()
        gb.blend_rect(a, src, 0, 0, mode)
        ob.blend_rect(b, src, 0, 0, mode)
        assert diff_report(a, b)[0] == 0, BLEND_MODE_NAMES[mode]


@pytest.mark.parametrize("radius", [1, 2, 3, 7, 20, 32, 65])
@pytest.mark.parametrize("shape", [(64, 64), (37, 101), (200, 13), (5, 300)])
def test_blur(radius, shape):
    gb, ob = _backends()
    h, w = shape
    img = synth.random_premultiplied(h, w, radius * 31 + w)
    lut = host.gaussianKernel(radius)
    for oob in (0, pack(10, 200, 30, 255), pack(5, 6, 7, 100)):
        a, b = img.copy(), img.copy()
        gb.blur(a, lut, radius, oob)
        ob.blur(b, lut, radius, oob)
        n, mx, where = diff_report(a, b)
        assert n == 0, f"r={radius} {shape} oob={oob:#x}: {n} px differ (max {mx}) at {where}"


def test_blur_large_radius_fallback_and_constant_image():
    gb, ob = _backends()
    img = synth.random_premultiplied(40, 40, 5)
    lut = host.gaussianKernel(720)  # above the tiled kernel's shared-memory limit
    a, b = img.copy(), img.copy()
    gb.blur(a, lut, 720, 0)
    ob.blur(b, lut, 720, 0)
    assert diff_report(a, b)[0] == 0
    # LUT sums to 65275 < 65280: a constant 255 image loses one LSB per pass (reference behaviour)
    c = np.full((128, 128, 4), 255, np.uint8)
    d = c.copy()
    gb.blur(c, host.gaussianKernel(32), 32, pack(255, 255, 255, 255))
    ob.blur(d, host.gaussianKernel(32), 32, pack(255, 255, 255, 255))
    assert diff_report(c, d)[0] == 0 and c.max() < 255


@pytest.mark.parametrize("amount", [1, 2, 5, -1, -3])
def test_spread(amount):
    gb, ob = _backends()
    img = synth.random_premultiplied(50, 70, 8)
    img[10:30, 20:50, 3] = 255
    a, b = img.copy(), img.copy()
    gb.spread(a, amount)
    ob.spread(b, amount)
    assert diff_report(a, b)[0] == 0


@pytest.mark.parametrize("offset", [(0, 0), (2, 2), (-7, 5), (40, -3)])
def test_shadow(offset):
    gb, ob = _backends()
    img = np.zeros((120, 160, 4), np.uint8)
    img[30:80, 40:110] = (200, 100, 50, 255)
    img[50:60, 60:70] = (10, 10, 10, 40)
    lut = host.gaussianKernel(10)
    a = gb.shadow(img, offset[0], offset[1], 4, lut, 10, pack(0, 0, 0, 200))
    b = ob.shadow(img, offset[0], offset[1], 4, lut, 10, pack(0, 0, 0, 200))
    assert diff_report(a, b)[0] == 0


def test_error_cases_match_reference():
    """PixieError conditions of the reference surface as status 1 + message."""
    from pixie_b200 import device as dev
    from pixie_b200.common import PixieError

    dev.init(0)
    with pytest.raises(PixieError, match="width and height must be > 0"):  # common.nim:41-42
        dev.DeviceImage(0, 10)
    img = dev.DeviceImage(16, 16)
    with pytest.raises(PixieError, match="negative blur"):  # images.nim:311-312
        dev.blur(img, host.gaussianKernel(1), -1)
    with pytest.raises(PixieError, match="different images"):
        dev.shadow(img, img, 0.5, 0, 1, host.gaussianKernel(2), 2, 0xFF000000)
    with pytest.raises(PixieError):
        dev.blend_rect(img, dev.DeviceImage(8, 8), 0, 0, 99)
    # huge coordinates: "Path int overflow detected" (paths.nim:1618-1619)
    segs = host.Segments(np.array([[-1e30, 0, 1e30, 10], [1e30, 0, -1e30, 10]], np.float32), np.array([1, -1], np.int16))
    with pytest.raises(PixieError, match="Path int overflow"):
        dev.fill_segments(img, segs, 0xFF0000FF, 0, 0)


@pytest.mark.parametrize("shape", [(37, 1030), (9, 2051), (130, 77), (5, 4), (64, 2048)])
@pytest.mark.parametrize("amount", [1, -1, 3, -6, 40, 2100])
def test_spread_shapes(shape, amount):
    """Row tiles of 1024 outputs, widths that are not multiples of 4, more rows than one CTA step, windows wider
    than the image, and the fallback beyond 2048."""
    gb, ob = _backends()
    img = synth.random_premultiplied(shape[0], shape[1], 9)
    a, b = img.copy(), img.copy()
    gb.spread(a, amount)
    ob.spread(b, amount)
    assert diff_report(a, b)[0] == 0


@pytest.mark.parametrize("offset", [(1030, 3), (-1500, -2), (5, 70), (3000, 0)])
@pytest.mark.parametrize("spread_", [2, -3])
def test_shadow_wide(offset, spread_):
    """The offset copy folded into the spread's read, on an image wider than one row tile (and offsets that push
    the source out of the image)."""
    gb, ob = _backends()
    img = synth.random_premultiplied(40, 2300, 10)
    lut = host.gaussianKernel(5)
    a = gb.shadow(img, offset[0], offset[1], spread_, lut, 5, pack(10, 20, 30, 200))
    b = ob.shadow(img, offset[0], offset[1], spread_, lut, 5, pack(10, 20, 30, 200))
    assert diff_report(a, b)[0] == 0


@pytest.mark.parametrize("shape", [(37, 130), (131, 77), (5, 4), (64, 256), (300, 515)])
@pytest.mark.parametrize("radius,spread_", [(1, 0), (3, 1), (4, -2), (8, 0), (29, 3), (32, 0), (64, 2), (70, 1), (0, 2), (0, 0)])
def test_shadow_alpha_plane(shape, radius, spread_):
    """shadow() with an integral offset runs offset copy, spread, blur and composite on the mask's alpha plane
    (one-channel tensor-core blur): every tile shape of that kernel — widths off the 4-byte grid, radii on and off
    the cp.async alignment, LUTs with and without high tap parts, no spread, no blur — and the RGBX fallback above
    radius 64."""
    gb, ob = _backends()
    img = synth.random_premultiplied(shape[0], shape[1], 11 + radius)
    img[: shape[0] // 3] = 0
    lut = host.gaussianKernel(radius)
    a = gb.shadow(img, 3, -2, spread_, lut, radius, pack(90, 20, 130, 220))
    b = ob.shadow(img, 3, -2, spread_, lut, radius, pack(90, 20, 130, 220))
    assert diff_report(a, b)[0] == 0


@pytest.mark.parametrize("radius", [29, 30, 31, 32])
@pytest.mark.parametrize("shape", [(33, 128), (70, 132), (131, 260), (64, 4), (1, 512), (300, 388)])
def test_blur_fused_tensor_core_path(radius, shape):
    """The fused tcgen05 kernel (blur_tc.cu: radii whose taps stay below 2048, widths that are multiples of 4): strips
    that end inside the last 128 columns, chunks that end inside a block of 32 rows, images shorter than one block,
    out-of-bounds colours (rows above / below the image enter the vertical pass as the colour itself)."""
    gb, ob = _backends()
    h, w = shape
    img = synth.random_premultiplied(h, w, radius * 7 + w)
    lut = host.gaussianKernel(radius)
    assert int(lut.max()) < 2048
    for oob in (0, pack(10, 200, 30, 255), pack(255, 255, 255, 255)):
        a, b = img.copy(), img.copy()
        gb.blur(a, lut, radius, oob)
        ob.blur(b, lut, radius, oob)
        n, mx, where = diff_report(a, b)
        assert n == 0, f"r={radius} {shape} oob={oob:#x}: {n} px differ (max {mx}) at {where}"


def test_blur_fused_path_moves_owned_images_and_keeps_wrapped_ones():
    """A whole-image blur of a library-owned image switches the handle to a fresh buffer (the fused pass cannot run in
    place); a caller-owned (wrapped) image keeps its memory and gets the rows copied back."""
    import torch

    from pixie_b200 import device as dev

    dev.init()
    _, ob = _backends()
    h, w, r = 200, 256, 32
    img = synth.random_premultiplied(h, w, 77)
    lut = host.gaussianKernel(r)
    want = img.copy()
    ob.blur(want, lut, r, 0)
    t = torch.from_numpy(img.copy()).cuda()
    wrapped = dev.DeviceImage.wrap(t.data_ptr(), w, h, owner=t)
    dev.blur(wrapped, lut, r, 0)
    dev.sync()
    assert wrapped.device_ptr() == t.data_ptr()
    assert diff_report(t.cpu().numpy(), want)[0] == 0
    owned = dev.DeviceImage(w, h).upload(img)
    dev.blur(owned, lut, r, 0)
    assert diff_report(owned.download(), want)[0] == 0
    dev.blur(owned, lut, r, 0)  # and again on the moved buffer
    ob.blur(want, lut, r, 0)
    assert diff_report(owned.download(), want)[0] == 0
