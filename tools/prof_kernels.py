"""Runs each hot kernel a few times so that ncu can capture it (see profiles/README.md)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixie_b200 import device as dev, host, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev.init(0)
if which in ("blend", "all"):
    n = 8192
    dst = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 1), (n // 512, 1, 1)))
    src = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 2), (n // 512, 1, 1)))
    mask = dev.DeviceImage(n, n, a8=True).upload(np.tile(synth.coverage_mask(512, n, 3), (n // 512, 1)))
    modes = [int(v) for v in os.environ.get("BLEND_MODES", "0,7,3,8,12,16").split(",")]
    for mode in modes:  # common.py ordinals: Normal 0, Overlay 7, ColorBurn 3, SoftLight 8, Hue 12, Mask 16
        for _ in range(2):
            dev.blend_rect_masked(dst, src, mask, 0, 0, mode)
    dev.sync()
if which in ("blur", "all"):
    n = int(os.environ.get("BLUR_N", "16384"))
    img = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 4), (n // 512, 1, 1)))
    for _ in range(3):
        dev.blur(img, host.gaussianKernel(32), 32, 0)
    dev.sync()
if which in ("shadow", "all"):
    n = int(os.environ.get("BLUR_N", "16384"))
    img = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 4), (n // 512, 1, 1)))
    out = dev.DeviceImage(n, n)
    for _ in range(3):
        dev.shadow(img, out, 8, 8, 4, host.gaussianKernel(32), 32, 0xC8000000)
    dev.sync()
if which in ("tiger", "all"):
    sys.path.insert(0, ROOT)
    from bench import tiger_arrays

    arrays = tiger_arrays(4096)
    img = dev.DeviceImage(4096, 4096)
    cl = dev.CmdList(4096, 4096, 1, arrays)
    for _ in range(3):
        cl.run(img, clear=True)  # as the bench step: the raster kernel clears the canvas
    dev.sync()
if which in ("icons", "all"):
    from pixie_b200.device import FillBatch

    batch = FillBatch()
    for i in range(1024):
        synth.icon_fills(i, 512, i, batch)
    arrays = batch.arrays()
    img = dev.DeviceImage(512, 512, 1024)
    cl = dev.CmdList(512, 512, 1024, arrays)
    for _ in range(3):
        img.fill(0)
        cl.run(img)
    dev.sync()
if which in ("flatten", "all"):
    from pixie_b200 import svg as psvg

    doc = psvg.parseSvg(open(os.path.join(ROOT, "tests", "golden", "tiger.svg")).read(), 4096, 4096)
    pb = psvg.svg_path_batch(doc)
    packed = pb.packed()
    for _ in range(3):
        cl = dev.CmdList.from_paths(4096, 4096, 1, pb, packed)
        del cl
    dev.sync()
if which in ("paint", "all"):
    n = 8192
    dst = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 1), (n // 512, 1, 1)))
    mask = dev.DeviceImage(n, n)
    ell = host.newPath()
    ell.ellipse(n / 2, n / 2, n * 0.3, n * 0.22)
    dev.fill_segments(mask, host.fill_segments(ell), 0xFFFFFFFF, 0, 0)
    stops = [(0.0, (1, 0, 0, 1)), (0.3, (0, 1, 0, 0.5)), (1.0, (0, 0, 1, 1))]
    for kind, handles in ((3, [(n * 0.2, n * 0.3), (n * 0.9, n * 0.7)]), (4, [(n / 2, n / 2), (n * 0.9, n / 2), (n / 2, n * 0.95)])):
        for _ in range(2):
            dev.fill_gradient_masked(dst, mask, kind, handles, stops, 1.0, 0)
    dev.sync()
if which in ("draw", "all"):
    n = 8192
    f = np.float32
    dst = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 1), (n // 512, 1, 1)))
    src = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 2), (n // 512, 1, 1)))
    rot = host.matmul(host.translate(f(n / 2), f(-n / 5)), host.rotate(f(0.5)))
    dev.draw(dst, src, rot, 0)
    dev.draw(dst, src, host.scale(f(0.5), f(0.5)), 0)
    dev.fill_gradient(dst, 4, [(n / 2, n / 2), (n, n / 2), (n / 2, n)], [(0.0, (1, 0, 0, 1)), (1.0, (0, 0, 1, 0.5))], 1.0)
    dev.sync()
