"""The reference's example programs written against pixie_b200.api (same proc names / arguments),
checked byte-for-byte against the reference's own output PNGs."""
import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu


def test_examples_heart():  # examples/heart.nim
    from pixie_b200 import api as pixie

    image = pixie.newImage(200, 200)
    image.fill((255, 255, 255, 255))
    image.fillPath(gc.HEART, "#FC427B")
    assert gc.compare(image.data, gc.load_golden("examples_heart.png")) == (0, 0)
    assert image[100, 100] == (252, 66, 123, 255) and image[-1, 5] == (0, 0, 0, 0)


def test_examples_shadow():  # examples/shadow.nim
    from pixie_b200 import api as pixie

    image = pixie.newImage(200, 200)
    image.fill((255, 255, 255, 255))
    path = pixie.newPath()
    path.polygon(100, 100, 70, 8)
    polygonImage = pixie.newImage(200, 200)
    polygonImage.fillPath(path, (255, 255, 255, 255))
    shadow = polygonImage.shadow(offset=(2, 2), spread=2, blur=10, color=(0, 0, 0, 200))
    image.draw(shadow)
    image.draw(polygonImage)
    assert gc.compare(image.data, gc.load_golden("examples_shadow.png")) == (0, 0)


def test_examples_blur_and_masking():  # examples/blur.nim
    from pixie_b200 import api as pixie
    from pixie_b200.common import MaskBlend

    trees = pixie.newImage(200, 200)
    trees.data = gc.load_golden("examples_data_trees.png")
    blur = trees.copy()
    image = pixie.newImage(200, 200)
    image.fill((255, 255, 255, 255))
    path = pixie.newPath()
    path.polygon(100, 100, 70, 6)
    mask = pixie.newImage(200, 200)
    mask.fillPath(path, (1.0, 1.0, 1.0, 1.0))
    blur.blur(20)
    blur.draw(mask, blendMode=MaskBlend)
    image.draw(trees)
    image.draw(blur)
    assert gc.compare(image.data, gc.load_golden("examples_blur.png")) == (0, 0)


def test_paint_opacity_stroke_and_errors():  # tests/test_paths.nim:595-606
    from pixie_b200 import api as pixie
    from pixie_b200.common import PixieError
    from pixie_b200 import host

    path = pixie.newPath()
    path.circle(50, 50, 30)
    paint = pixie.newPaint(pixie.SolidPaint)
    paint.color = (1.0, 0.0, 1.0, 1.0)
    paint.opacity = 0.5
    image = pixie.newImage(100, 100)
    image.strokePath(path, paint, strokeWidth=10)
    assert gc.compare(image.data, gc.load_golden("paths_opacityStroke.png")) == (0, 0)
    with pytest.raises(PixieError):
        pixie.newImage(0, 5)
    with pytest.raises(PixieError, match="negative blur"):
        image.blur(-3)
    with pytest.raises(PixieError, match="Cannot minifyBy2 with negative power"):
        image.minifyBy2(-1)


def test_image_paint_composite_matches_two_draws():
    """Non-solid (image) paint: mask + fill + fused masked draw == the reference's two draws (paths.nim:2115-2142)."""
    from pixie_b200 import api as pixie, synth
    from pixie_b200.common import MaskBlend, NormalBlend
    from _oracle import OracleBackend
    from pixie_b200 import host

    w = h = 128
    tex = synth.random_premultiplied(h, w, 77)
    bg = synth.random_premultiplied(h, w, 78)
    image = pixie.newImage(w, h)
    image.data = bg
    timg = pixie.newImage(w, h)
    timg.data = tex
    paint = pixie.newPaint(pixie.ImagePaint)
    paint.image = timg
    paint.opacity = 0.7
    path = "M 20.5 10 L 110 40.25 L 60 120 z"
    image.fillPath(path, paint)
    # reference sequence on the oracle
    ob = OracleBackend(0)
    mask = np.zeros((h, w, 4), np.uint8)
    ob.fill_segments(mask, host.fill_segments(path), 0xFFFFFFFF, 0, NormalBlend)
    fill = np.zeros((h, w, 4), np.uint8)
    ob.blend_rect(fill, tex, 0, 0, NormalBlend)
    from _oracle import lib as olib
    olib().orc_apply_opacity(mask.ctypes.data, w, h, 0.7)
    ob.blend_rect(fill, mask, 0, 0, MaskBlend)
    want = bg.copy()
    ob.blend_rect(want, fill, 0, 0, NormalBlend)
    assert gc.compare(image.data, want) == (0, 0)


def test_icons_batch_sharded_checksum():
    """BASELINE config 5 in miniature: 64 synthetic icons as 64 layers in one launch pair; results stay on
    the device, parity through the checksum-of-layers and a downloaded sample."""
    from pixie_b200 import device as dev, synth
    from pixie_b200.device import FillBatch
    from _util import oracle_render_batch

    dev.init(0)
    n, size = 64, 128
    b = FillBatch()
    for i in range(n):
        synth.icon_fills(5000 + i, size, i, b)
    arr = b.arrays()
    img = dev.DeviceImage(size, size, n)
    covered = dev.fill_batch(img, arr, count_covered=True)
    want, wc = oracle_render_batch(arr, size, size, layers=n)
    assert covered == wc
    flat = want.reshape(-1, 4).view(np.uint32).reshape(-1).astype(np.uint64)
    idx = np.arange(flat.size, dtype=np.uint64) | np.uint64(1)
    assert img.checksum() == int((flat * idx).sum(dtype=np.uint64))
    assert np.array_equal(img.download()[7], want[7])
