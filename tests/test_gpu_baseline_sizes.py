"""Parity against the ORACLE at BASELINE.json's full sizes (configs 3 and 4): blends at 8192^2, blur r=32 and the drop
shadow at 16384^2 — through windows (tests/_windows.py; the method itself is checked on the CPU in
tests/test_windows_cpu.py).  A 16384^2 canvas is 1 GiB: index arithmetic near 2^30-2^32 bytes and the persistent tile
schedulers at full grid size are what these cases add over the small-shape tests."""
import numpy as np
import pytest

from pixie_b200 import host, synth
from pixie_b200.common import BLEND_MODE_NAMES, rgbx as pack
import _windows as W

pytestmark = pytest.mark.gpu


def _tiled(tile_fn, n, seed):
    t = tile_fn(512, n, seed)
    return np.tile(t, (n // 512,) + (1,) * (t.ndim - 1))


@pytest.fixture(scope="module")
def blend_inputs():
    from pixie_b200 import device as dev

    dev.init()
    n = 8192
    dst = _tiled(synth.random_premultiplied, n, 0x5EED)
    src = _tiled(synth.random_premultiplied, n, 0x5EED + 1)
    mask = _tiled(synth.coverage_mask, n, 0x5EED + 2)
    # rows differ from tile to tile so that a row mix-up between tiles cannot hide
    dst[:, :, 0] ^= (np.arange(n, dtype=np.uint32) // 512 * 7 % 256).astype(np.uint8)[:, None]
    dst[..., :3] = np.minimum(dst[..., :3], dst[..., 3:4])
    d0 = dev.DeviceImage(n, n).upload(dst)
    s = dev.DeviceImage(n, n).upload(src)
    m = dev.DeviceImage(n, n, a8=True).upload(mask)
    return dev, n, dst, src, mask, d0, s, m


@pytest.mark.parametrize("mode", [0, 16, 7, 12, 3, 17], ids=lambda m: BLEND_MODE_NAMES[m])
def test_blend_8192_against_oracle(blend_inputs, mode):
    dev, n, dst, src, mask, d0, s, m = blend_inputs
    d = dev.DeviceImage(n, n)
    rows = [(0, 8), (509, 517), (4093, 4101), (n - 8, n)]
    d.copy_from(d0)
    dev.blend_rect_masked(d, s, m, 0, 0, mode)
    cnt, bad, mx = W.check_blend_rows(d, dst, src, mask, mode, rows)
    assert cnt == 32 * n and bad == 0, f"{BLEND_MODE_NAMES[mode]} masked: {bad} px differ (max {mx})"
    d.copy_from(d0)
    dev.blend_rect(d, s, 0, 0, mode)
    cnt, bad, mx = W.check_blend_rows(d, dst, src, None, mode, rows)
    assert bad == 0, f"{BLEND_MODE_NAMES[mode]}: {bad} px differ (max {mx})"


@pytest.fixture(scope="module")
def big_image():
    from pixie_b200 import device as dev

    dev.init()
    n = 16384
    img = _tiled(synth.random_premultiplied, n, 0xB10B)
    img[:, :, 1] ^= (np.arange(n, dtype=np.uint32) // 512 * 5 % 256).astype(np.uint8)[:, None]
    img[..., :3] = np.minimum(img[..., :3], img[..., 3:4])
    return dev, n, img


def test_blur_r32_16384_against_oracle(big_image):
    dev, n, img = big_image
    lut = host.gaussianKernel(32)
    d = dev.DeviceImage(n, n).upload(img)
    dev.blur(d, lut, 32, 0)
    wins = W.corner_and_seam_windows(n, n)
    cnt, bad, mx = W.check_blur_windows(d, img, lut, 32, 0, wins)
    assert cnt >= 7 * 48 * 48 and bad == 0, f"{bad} of {cnt} px differ (max {mx})"
    # out-of-bounds colour (images.nim:313-318) at the same size: the border windows see it
    d.upload(img)
    oob = pack(40, 30, 20, 120)
    dev.blur(d, lut, 32, oob)
    cnt, bad, mx = W.check_blur_windows(d, img, lut, 32, oob, wins[:4])
    assert bad == 0, f"oob: {bad} of {cnt} px differ (max {mx})"


def test_blur_structured_rect_16384(big_image):
    """SURVEY 8(d) C4's structured case: an opaque rectangle on a transparent canvas; windows on its corners."""
    dev, n, _ = big_image
    img = np.zeros((n, n, 4), np.uint8)
    img[3000:9000, 5000:14000] = (200, 100, 50, 255)
    lut = host.gaussianKernel(32)
    d = dev.DeviceImage(n, n).upload(img)
    dev.blur(d, lut, 32, 0)
    wins = [(2976, 3024, 4976, 5024), (8976, 9024, 13976, 14024), (2976, 3024, 9000, 9048), (6000, 6048, 4976, 5024),
            (0, 48, 0, 48), (n - 48, n, n - 48, n)]
    cnt, bad, mx = W.check_blur_windows(d, img, lut, 32, 0, wins)
    assert bad == 0, f"{bad} of {cnt} px differ (max {mx})"


def test_shadow_16384_against_oracle(big_image):
    dev, n, img = big_image
    lut = host.gaussianKernel(32)
    src = img.copy()
    src[: n // 3] = 0  # a shadow needs transparent surroundings to show
    s = dev.DeviceImage(n, n).upload(src)
    d = dev.DeviceImage(n, n)
    col = pack(0, 0, 0, 200)
    dev.shadow(s, d, 8.0, 8.0, 4, lut, 32, col)
    wins = W.corner_and_seam_windows(n, n, seams=((n // 3, 4096), (8192, 8192 + 32)))
    cnt, bad, mx = W.check_shadow_windows(d, src, (8, 8), 4, lut, 32, col, wins)
    assert cnt >= 7 * 48 * 48 and bad == 0, f"{bad} of {cnt} px differ (max {mx})"


def test_spread_16384_against_oracle(big_image):
    dev, n, img = big_image
    d = dev.DeviceImage(n, n).upload(img)
    dev.spread(d, 4)
    cnt, bad, mx = W.check_spread_windows(d, img, 4, W.corner_and_seam_windows(n, n, seams=((8192, 1024), (2048, 4096))))
    assert bad == 0, f"{bad} of {cnt} px differ (max {mx})"
