"""The reference-facing boundary pieces: every *_host entry point the Nim shim binds (host pixels in, host pixels out),
the row-band entry points, stream / device switching, and calls from several host threads."""
import ctypes as C
import threading

import numpy as np
import pytest

from pixie_b200 import host, synth
from pixie_b200.common import BLEND_MODE_NAMES, MaskBlend, NormalBlend, PixieError, rgbx as pack
from _util import diff_report

pytestmark = pytest.mark.gpu


def _o():
    from _oracle import OracleBackend

    return OracleBackend(0)


def _lib():
    from pixie_b200 import device as dev

    dev.init()
    return dev.lib(), dev


@pytest.mark.parametrize("mode", [0, 3, 7, 14, 16, 17, 19], ids=lambda m: BLEND_MODE_NAMES[m])
def test_blend_rect_host(mode):
    L, dev = _lib()
    dst = synth.random_premultiplied(70, 131, 5)
    src = synth.random_premultiplied(50, 90, 6)
    a, b = dst.copy(), dst.copy()
    dev.check(L.pixie_cuda_blend_rect_host(a.ctypes.data, 131, 70, src.ctypes.data, 90, 50, 17, -9, mode))
    _o().blend_rect(b, src, 17, -9, mode)
    assert diff_report(a, b)[0] == 0


@pytest.mark.parametrize("radius,oob", [(20, 0), (32, pack(9, 8, 7, 200)), (70, 0)])
def test_blur_host(radius, oob):
    L, dev = _lib()
    img = synth.random_premultiplied(90, 150, radius)
    lut = np.ascontiguousarray(host.gaussianKernel(radius), np.uint16)
    a, b = img.copy(), img.copy()
    dev.check(L.pixie_cuda_blur_host(a.ctypes.data, 150, 90, lut.ctypes.data, radius, oob))
    _o().blur(b, lut, radius, oob)
    assert diff_report(a, b)[0] == 0


@pytest.mark.parametrize("offset", [(3.0, -2.0), (2.5, 1.25)])
def test_shadow_host(offset):
    L, dev = _lib()
    img = np.zeros((100, 140, 4), np.uint8)
    img[30:70, 40:100] = (200, 100, 50, 255)
    lut = np.ascontiguousarray(host.gaussianKernel(8), np.uint16)
    out = np.full_like(img, 77)  # every byte of dst is produced by the call
    dev.check(L.pixie_cuda_shadow_host(img.ctypes.data, out.ctypes.data, 140, 100, offset[0], offset[1], 3, lut.ctypes.data, 8,
                                       pack(0, 0, 0, 200)))
    want = _o().shadow(img, offset[0], offset[1], 3, lut, 8, pack(0, 0, 0, 200))
    assert diff_report(out, want)[0] == 0


def test_spread_apply_opacity_masked_host():
    L, dev = _lib()
    img = synth.random_premultiplied(60, 77, 2)
    a, b = img.copy(), img.copy()
    dev.check(L.pixie_cuda_spread_host(a.ctypes.data, 77, 60, -3))
    _o().spread(b, -3)
    assert diff_report(a, b)[0] == 0
    a, b = img.copy(), img.copy()
    dev.check(L.pixie_cuda_apply_opacity_host(a.ctypes.data, 77, 60, 0.4))
    _o().apply_opacity(b, 0.4)
    assert diff_report(a, b)[0] == 0
    src = synth.random_premultiplied(40, 50, 3)
    for mask in (synth.coverage_mask(40, 50, 4), synth.random_premultiplied(40, 50, 5)):
        a, b = img.copy(), img.copy()
        dev.check(L.pixie_cuda_blend_rect_masked_host(a.ctypes.data, 77, 60, src.ctypes.data, mask.ctypes.data,
                                                      1 if mask.ndim == 2 else 4, 50, 40, 9, 11, NormalBlend))
        _o().blend_rect_masked(b, src, mask, 9, 11, NormalBlend)
        assert diff_report(a, b)[0] == 0
    with pytest.raises(PixieError, match="mask_bytes_per_pixel"):
        dev.check(L.pixie_cuda_blend_rect_masked_host(a.ctypes.data, 77, 60, src.ctypes.data, src.ctypes.data, 2, 50, 40, 0, 0, 0))


def test_blur_rows_every_band_against_the_oracle():
    """Row-band blur, every band compared with the ORACLE's whole-image blur: border bands (true image border on one
    side, halo on the other), an interior band, and the split X / Y entry points the overlapped exchange uses."""
    _, dev = _lib()
    h, w, r = 300, 128, 32
    img = synth.random_premultiplied(h, w, 42)
    lut = host.gaussianKernel(r)
    oob = pack(30, 20, 10, 90)
    want = img.copy()
    _o().blur(want, lut, r, oob)
    for (y0, y1) in [(0, 100), (100, 200), (200, 300), (0, 40), (260, 300)]:
        e0, e1 = max(0, y0 - r), min(h, y1 + r)
        ext = np.ascontiguousarray(img[e0:e1])
        # halo rows present wherever the image continues; an edge of ext without halo is the image's own border
        d = dev.DeviceImage(w, e1 - e0).upload(ext)
        dev.blur_rows(d, lut, r, oob, y0 - e0, y1 - e0)
        got = d.download()
        n, mx, where = diff_report(got[y0 - e0:y1 - e0], want[y0:y1])
        assert n == 0, f"band {(y0, y1)}: {n} px differ (max {mx}) at {where}"
        assert np.array_equal(got[:y0 - e0], ext[:y0 - e0]) and np.array_equal(got[y1 - e0:], ext[y1 - e0:]), \
            "rows outside [y0, y1) must be left alone"
        # the same band through the split passes: X on the band rows, then X on the halo rows, then Y
        d2 = dev.DeviceImage(w, e1 - e0).upload(ext)
        dev.blur_rows_x(d2, lut, r, oob, y0 - e0, y1 - e0)
        if y0 - e0:
            dev.blur_rows_x(d2, lut, r, oob, 0, y0 - e0)
        if e1 - y1:
            dev.blur_rows_x(d2, lut, r, oob, y1 - e0, e1 - e0)
        dev.blur_rows_y(d2, lut, r, oob, y0 - e0, y1 - e0)
        assert diff_report(d2.download()[y0 - e0:y1 - e0], want[y0:y1])[0] == 0, f"split passes, band {(y0, y1)}"


@pytest.mark.parametrize("amount", [5, -4])
def test_spread_rows_bands_against_the_oracle(amount):
    _, dev = _lib()
    h, w = 200, 96
    img = synth.random_premultiplied(h, w, 7)
    want = img.copy()
    _o().spread(want, amount)
    s = abs(amount)
    for (y0, y1) in [(0, 70), (70, 140), (140, 200)]:
        e0, e1 = max(0, y0 - s), min(h, y1 + s)
        d = dev.DeviceImage(w, e1 - e0).upload(np.ascontiguousarray(img[e0:e1]))
        dev.spread_rows(d, amount, y0 - e0, y1 - e0)
        assert diff_report(d.download()[y0 - e0:y1 - e0], want[y0:y1])[0] == 0, (y0, y1)
    d = dev.DeviceImage(w, 50).upload(np.ascontiguousarray(img[10:60]))
    with pytest.raises(PixieError, match="halo shorter"):
        dev.spread_rows(d, amount, 2, 40)


def test_shadow_rows_bands_against_the_oracle():
    _, dev = _lib()
    h, w, r, sp, off = 260, 160, 12, 3, (5.0, -6.0)
    img = synth.random_premultiplied(h, w, 8)
    img[:60] = 0
    lut = host.gaussianKernel(r)
    col = pack(10, 20, 30, 200)
    want = _o().shadow(img, off[0], off[1], sp, lut, r, col)
    need = 6 + sp + r
    for (y0, y1) in [(0, 90), (90, 170), (170, 260)]:
        e0, e1 = max(0, y0 - need), min(h, y1 + need)
        s = dev.DeviceImage(w, e1 - e0).upload(np.ascontiguousarray(img[e0:e1]))
        d = dev.DeviceImage(w, e1 - e0)
        dev.shadow_rows(s, d, off[0], off[1], sp, lut, r, col, y0 - e0, y1 - e0)
        n, mx, where = diff_report(d.download()[y0 - e0:y1 - e0], want[y0:y1])
        assert n == 0, f"band {(y0, y1)}: {n} px differ (max {mx}) at {where}"


def test_cmdlist_run_rows_equals_whole_canvas():
    """A canvas of fills rendered band by band (pixie_cuda_cmdlist_run_rows, band images and in-place rows) equals
    the undivided render and the oracle — including MaskBlend fills whose clears reach other rows and whose
    negative-x clears read the plans of later rows across a band edge (mask_wrap_clears)."""
    _, dev = _lib()
    from _util import oracle_render_batch

    size = 384
    batch = dev.FillBatch()
    for i in range(6):
        synth.icon_fills(100 + i, size, 0, batch)
    # a MaskBlend path that crosses x < 0 (its clears wrap into earlier rows) and several band edges
    p = host.newPath()
    p.moveTo(-700.5, 20.25)
    p.lineTo(300.0, 60.5)
    p.lineTo(250.75, 330.0)
    p.lineTo(-650.0, 300.5)
    p.closePath()
    batch.add(host.fill_segments(p), pack(255, 255, 255, 255), 0, MaskBlend)
    synth.icon_fills(200, size, 0, batch)
    arrays = batch.arrays()
    want, covered = oracle_render_batch(arrays, size, size)
    cl = dev.CmdList(size, size, 1, arrays)
    whole = dev.DeviceImage(size, size)
    cov_whole = cl.run(whole, count_covered=True)
    assert diff_report(whole.download(), want[0])[0] == 0 and cov_whole == covered
    edges = [0, 97, 192, 193, 300, size]
    inplace = dev.DeviceImage(size, size)
    got = np.zeros((size, size, 4), np.uint8)
    cov = 0
    for y0, y1 in zip(edges[:-1], edges[1:]):
        band = dev.DeviceImage(size, y1 - y0)
        cov += cl.run_rows(band, y0, y1, count_covered=True)
        got[y0:y1] = band.download()
        cl.run_rows(inplace, y0, y1)
    n, mx, where = diff_report(got, want[0])
    assert n == 0, f"band images: {n} px differ (max {mx}) at {where}"
    assert cov == covered
    assert diff_report(inplace.download(), want[0])[0] == 0


def test_render_batch_host_with_wrapping_mask_fill():
    """a MaskBlend fill with xMin < 0 makes a row read the plans of later rows; the banded host render
    must not let a band's raster kernel run ahead of the next band's plan kernels."""
    _, dev = _lib()
    from _util import oracle_render_batch

    size = 512
    batch = dev.FillBatch()
    for i in range(4):
        synth.icon_fills(300 + i, size, 0, batch)
    p = host.newPath()
    p.moveTo(-1500.25, 10.5)
    p.lineTo(400.0, 30.0)
    p.lineTo(380.5, 500.0)
    p.lineTo(-1400.0, 480.25)
    p.closePath()
    batch.add(host.fill_segments(p), pack(255, 255, 255, 255), 0, MaskBlend)
    arrays = batch.arrays()
    want, _ = oracle_render_batch(arrays, size, size)
    out = np.zeros((size, size, 4), np.uint8)
    for _ in range(5):
        out[:] = 0x55
        dev.render_batch_host(out.ctypes.data, size, size, arrays)
        assert diff_report(out, want[0])[0] == 0


def test_second_device_is_rejected_and_stream_switch_is_ordered():
    L, dev = _lib()
    cur = dev.current_device()
    with pytest.raises(PixieError, match="one process drives one GPU"):
        dev.check(L.pixie_cuda_init(cur + 1 if dev.device_count() > cur + 1 else cur + 1))
    dev.init(cur)  # same device: idempotent
    # images created on the library stream, used after switching to another stream and back
    import torch

    img = dev.DeviceImage(256, 256)
    img.fill(0x80402010)
    st = torch.cuda.Stream()
    dev.set_stream(st.cuda_stream)
    try:
        lut = host.gaussianKernel(3)
        dev.blur(img, lut, 3, 0x80402010)
    finally:
        dev.set_stream(None)
    got = img.download()
    want = np.empty((256, 256, 4), np.uint8)
    want[:] = (0x10, 0x20, 0x40, 0x80)
    _o().blur(want, lut, 3, 0x80402010)
    assert diff_report(got, want)[0] == 0


def test_two_host_threads_on_distinct_handles():
    """include/pixie_cuda.h "threading": one coarse lock — calls from two threads on distinct images are safe."""
    _, dev = _lib()
    lut = host.gaussianKernel(5)
    results, errors = {}, []

    def work(seed):
        try:
            img = synth.random_premultiplied(96, 160, seed)
            src = synth.random_premultiplied(96, 160, seed + 50)
            want = img.copy()
            o = _o()
            for k in range(6):
                o.blend_rect(want, src, k, -k, NormalBlend)
                o.blur(want, lut, 5, 0)
            d = dev.DeviceImage(160, 96).upload(img)
            s = dev.DeviceImage(160, 96).upload(src)
            for k in range(6):
                dev.blend_rect(d, s, k, -k, NormalBlend)
                dev.blur(d, lut, 5, 0)
            results[seed] = diff_report(d.download(), want)[0]
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    ts = [threading.Thread(target=work, args=(s,)) for s in (1, 2, 3, 4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    assert results == {1: 0, 2: 0, 3: 0, 4: 0}


@pytest.mark.parametrize("radius,w", [(32, 128), (30, 260), (20, 96), (32, 101), (40, 64)])
def test_blur_rows_to_composes_interior_and_edges(radius, w):
    """pixie_cuda_blur_rows_to: out of place, src untouched, dst rows outside the range untouched; a band blurred as
    interior rows + two edge strips (how multi.RowBand overlaps the halo exchange) equals the oracle — on the fused
    tcgen05 path (taps < 2048, width % 4 == 0) and on the fallback (other radii / widths)."""
    _, dev = _lib()
    h = 260
    img = synth.random_premultiplied(h, w, radius + w)
    lut = host.gaussianKernel(radius)
    oob = pack(7, 9, 11, 130)
    want = img.copy()
    _o().blur(want, lut, radius, oob)
    s = dev.DeviceImage(w, h).upload(img)
    canary = np.full((h, w, 4), 0x5A, np.uint8)
    d = dev.DeviceImage(w, h).upload(canary)
    y0, y1 = 40, 230
    dev.blur_rows_to(s, d, lut, radius, oob, y0 + radius, y1 - radius)
    dev.blur_rows_to(s, d, lut, radius, oob, y0, y0 + radius)
    dev.blur_rows_to(s, d, lut, radius, oob, y1 - radius, y1)
    got = d.download()
    n, mx, where = diff_report(got[y0:y1], want[y0:y1])
    assert n == 0, f"{n} px differ (max {mx}) at {where}"
    assert np.array_equal(got[:y0], canary[:y0]) and np.array_equal(got[y1:], canary[y1:]), "rows outside [y0, y1) of dst changed"
    assert np.array_equal(s.download(), img), "src changed"
