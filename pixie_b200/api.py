"""Pixie's public procs for the raster hot path, on device-resident images.

Same names, argument meaning and error behaviour as the reference (treeform/pixie):
``newImage`` (common.nim:39-47), ``Image.fill`` (pixie.nim:120-131), ``fillPath`` / ``strokePath``
(paths.nim:2093-2214), ``draw`` (images.nim:636-678, integer-translate path = blendRect),
``blur`` (images.nim:304-365), ``shadow`` (images.nim:760-776), ``applyOpacity`` (images.nim:261-277),
``Paint`` (paints.nim:12-24).  The host part (path parsing, flattening, stroking, segment
building, Gaussian LUT) runs in libpixie_host.so as it does in Nim; everything that touches pixels
runs in pixie_cuda.so.  There is no CPU fallback.

``draw`` takes any transform (minifyBy2 / magnifyBy2 chain + drawSmooth, images.nim:531-678), ``Paint`` covers
all six PaintKinds (paints.nim:4-24): image / tiled image / linear, radial, angular gradients are composited as
the reference does (paths.nim:2115-2142) with the last two draws fused into one masked blend.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import device as dev
from . import host
from .common import (MaskBlend, NormalBlend, OverwriteBlend, PixieError, parseHtmlColor, rgbx as pack_rgbx)
from .host import (BevelJoin, ButtCap, EvenOdd, MiterJoin, NonZero, RoundCap, RoundJoin, SquareCap,  # noqa: F401
                   Path, defaultMiterLimit, newPath, parsePath)

# paints.nim:4-10 PaintKind
SolidPaint, ImagePaint, TiledImagePaint, LinearGradientPaint, RadialGradientPaint, AngularGradientPaint = range(6)


def _color_to_rgbx(color, opacity=1.0) -> int:
    """chroma Color (straight float r,g,b,a) -> color.a *= opacity -> asRgbx() (paths.nim:2110-2112,1603)."""
    f = np.float32
    r, g, b, a = (f(v) for v in color)
    a = a * f(opacity)

    def q(v):
        return max(0, min(255, int(math.floor(float(v * f(255)) + 0.5))))

    r8, g8, b8, a8 = q(r), q(g), q(b), q(a)
    if a8 != 255:
        r8, g8, b8 = (r8 * a8 + 127) // 255, (g8 * a8 + 127) // 255, (b8 * a8 + 127) // 255
    return pack_rgbx(r8, g8, b8, a8)


def _some_color(c):
    """SomeColor: '#rrggbb' string, (r,g,b,a) floats in [0,1], or ints 0..255 (ColorRGBA)."""
    if isinstance(c, str):
        r, g, b, a = parseHtmlColor(c)
        return (r / 255.0, g / 255.0, b / 255.0, a / 255.0)
    c = tuple(c)
    if all(isinstance(v, (int, np.integer)) for v in c):
        return tuple(np.float32(v) / np.float32(255) for v in c)
    return c


@dataclass
class Paint:
    """paints.nim:12-24."""
    kind: int = SolidPaint
    blendMode: int = NormalBlend
    opacity: float = 1.0
    color: tuple = (0.0, 0.0, 0.0, 1.0)
    image: "Image | None" = None
    imageMat: np.ndarray = field(default_factory=host.mat3)
    gradientHandlePositions: list = field(default_factory=list)  # [(x, y)] in image space
    gradientStops: list = field(default_factory=list)            # [ColorStop]


@dataclass
class ColorStop:
    """paints.nim:26-29."""
    color: tuple = (0.0, 0.0, 0.0, 1.0)
    position: float = 0.0


def newPaint(kind=SolidPaint) -> Paint:
    return Paint(kind=kind)


def _some_paint(p) -> Paint:
    if isinstance(p, Paint):
        return p
    return Paint(color=_some_color(p))


class Image:
    """common.nim:34-37 Image, premultiplied RGBX, resident in HBM."""

    def __init__(self, width, height, _dev=None):
        dev.init(dev_index())
        self._d = _dev if _dev is not None else dev.DeviceImage(width, height)
        self.width, self.height = width, height

    # -- host views
    @property
    def data(self) -> np.ndarray:
        return self._d.download()

    @data.setter
    def data(self, pixels):
        self._d.upload(pixels)

    def copy(self) -> "Image":
        out = Image(self.width, self.height)
        out._d.copy_from(self._d)
        return out

    def __getitem__(self, xy):
        x, y = xy
        if x < 0 or y < 0 or x >= self.width or y >= self.height:
            return (0, 0, 0, 0)
        return tuple(int(v) for v in self._d.download_rows(y, y + 1)[0, x])

    # -- pixie.nim:120-131
    def fill(self, color):
        if isinstance(color, Paint):
            paint = color
            if paint.kind == SolidPaint:
                self._d.fill(_color_to_rgbx(paint.color))
            elif paint.kind in (ImagePaint, TiledImagePaint):
                self._d.fill(0)
                path = newPath()
                path.rect(0, 0, float(self.width), float(self.height))
                self.fillPath(path, paint)
            else:
                self.fillGradient(paint)
        else:
            self._d.fill(_color_to_rgbx(_some_color(color)))

    # -- paints.nim:236-248
    def fillGradient(self, paint: Paint):
        if paint.kind not in (LinearGradientPaint, RadialGradientPaint, AngularGradientPaint):
            raise PixieError("Paint must be a gradient")
        # ColorStop.color is a chroma Color: straight float r, g, b, a (or an HTML colour string)
        stops = [(float(st.position), tuple(float(v) for v in (_some_color(st.color) if isinstance(st.color, str) else st.color)))
                 for st in paint.gradientStops]
        dev.fill_gradient(self._d, paint.kind, [tuple(h) for h in paint.gradientHandlePositions], stops, paint.opacity)

    # -- paths.nim:2093-2142
    def fillPath(self, path, paint, transform=None, windingRule=NonZero):
        paint = _some_paint(paint)
        paint.opacity = min(max(paint.opacity, 0.0), 1.0)
        if paint.opacity == 0:
            return
        if paint.kind == SolidPaint:
            if paint.color[3] > 0 or paint.blendMode == OverwriteBlend:
                segs = host.fill_segments(path, transform)
                dev.fill_segments(self._d, segs, _color_to_rgbx(paint.color, paint.opacity), windingRule, paint.blendMode)
            return
        mask = Image(self.width, self.height)
        mask.fillPath(path, (1.0, 1.0, 1.0, 1.0), transform, windingRule)
        self._composite_non_solid(mask, paint)

    # -- paths.nim:2144-2214
    def strokePath(self, path, paint, transform=None, strokeWidth=1.0, lineCap=ButtCap, lineJoin=MiterJoin,
                   miterLimit=defaultMiterLimit, dashes=()):
        paint = _some_paint(paint)
        paint.opacity = min(max(paint.opacity, 0.0), 1.0)
        if paint.opacity == 0:
            return
        if paint.kind == SolidPaint:
            if paint.color[3] > 0 or paint.blendMode == OverwriteBlend:
                segs = host.stroke_segments(path, transform, strokeWidth, lineCap, lineJoin, miterLimit, dashes)
                dev.fill_segments(self._d, segs, _color_to_rgbx(paint.color, paint.opacity), NonZero, paint.blendMode)
            return
        mask = Image(self.width, self.height)
        mask.strokePath(path, (1.0, 1.0, 1.0, 1.0), transform, strokeWidth, lineCap, lineJoin, miterLimit, dashes)
        self._composite_non_solid(mask, paint)

    def _composite_non_solid(self, mask: "Image", paint: Paint):
        """paths.nim:2115-2142: fill image + mask, `fill.draw(mask, MaskBlend); image.draw(fill, blendMode)`
        fused into one pass (pixie_cuda_blend_rect_masked); a gradient paint is evaluated inside that pass
        (pixie_cuda_fill_gradient_masked): no fill image at all."""
        if paint.kind in (LinearGradientPaint, RadialGradientPaint, AngularGradientPaint):
            if len(paint.gradientStops) == 0:
                raise PixieError("Gradient must have at least 1 color stop")
            stops = [(float(st.position), tuple(float(v) for v in (_some_color(st.color) if isinstance(st.color, str) else st.color)))
                     for st in paint.gradientStops]
            dev.fill_gradient_masked(self._d, mask._d, paint.kind, [tuple(h) for h in paint.gradientHandlePositions], stops,
                                     paint.opacity, paint.blendMode)
            return
        fill = Image(self.width, self.height)
        if paint.kind == ImagePaint:
            dev.draw(fill._d, paint.image._d, paint.imageMat, NormalBlend)        # fill.draw(paint.image, paint.imageMat)
        elif paint.kind == TiledImagePaint:
            dev.draw_tiled(fill._d, paint.image._d, paint.imageMat, NormalBlend)  # fill.drawTiled(...)
        else:
            saved, paint.opacity = paint.opacity, 1.0                             # paths.nim:2123-2136
            try:
                fill.fillGradient(paint)
            finally:
                paint.opacity = saved
        if paint.opacity != 1:
            dev.apply_opacity(mask._d, paint.opacity)
        dev.blend_rect_masked(self._d, fill._d, mask._d, 0, 0, paint.blendMode)

    # -- images.nim:636-678
    def draw(self, other: "Image", transform=None, blendMode=NormalBlend):
        dev.draw(self._d, other._d, host.mat3() if transform is None else transform, blendMode)

    def drawTiled(self, other: "Image", mat, blendMode=NormalBlend):
        dev.draw_tiled(self._d, other._d, mat, blendMode)

    # -- images.nim:168-259
    def minifyBy2(self, power=1) -> "Image":
        d = dev.minify_by2(self._d, power)
        return Image(d.width, d.height, _dev=d)

    def magnifyBy2(self, power=1) -> "Image":
        d = dev.magnify_by2(self._d, power)
        return Image(d.width, d.height, _dev=d)

    # -- images.nim:685-698
    def resize(self, width, height) -> "Image":
        if width == self.width and height == self.height:
            return self.copy()
        out = newImage(width, height)
        f = np.float32
        out.draw(self, host.scale(f(width) / f(self.width), f(height) / f(self.height)), OverwriteBlend)
        return out

    def applyOpacity(self, opacity):
        dev.apply_opacity(self._d, opacity)

    # -- images.nim:304-365
    def blur(self, radius, outOfBounds=(0.0, 0.0, 0.0, 0.0)):
        r = int(math.floor(radius + 0.5)) if radius >= 0 else -int(math.floor(-radius + 0.5))
        if r == 0:
            return
        if r < 0:
            raise PixieError("Cannot apply negative blur")
        dev.blur(self._d, host.gaussianKernel(r), r, _color_to_rgbx(_some_color(outOfBounds)))

    # -- images.nim:760-776
    def shadow(self, offset, spread, blur, color) -> "Image":
        out = Image(self.width, self.height)
        sp = int(math.floor(abs(spread) + 0.5)) * (1 if spread >= 0 else -1)
        r = int(math.floor(blur + 0.5))
        if r < 0:
            raise PixieError("Cannot apply negative blur")
        lut = host.gaussianKernel(max(r, 0))
        dev.shadow(self._d, out._d, float(offset[0]), float(offset[1]), sp, lut, r, _color_to_rgbx(_some_color(color)))
        return out


_DEVICE = None


def dev_index():
    """The device api.Image uses: the one set_device() / device.init() selected, else 0."""
    return _DEVICE if _DEVICE is not None else dev.current_device()


def set_device(index: int):
    """One process per GPU: select the device before the first Image is created."""
    global _DEVICE
    _DEVICE = index
    dev.init(index)


def newImage(width, height) -> Image:
    if width <= 0 or height <= 0:
        raise PixieError("Image width and height must be > 0")
    return Image(width, height)


def readSvg(data: str, width=0, height=0) -> Image:
    """newImage(parseSvg(data, width, height)) for the path-only SVG subset (svg.nim:502-608): one
    ordered command list, one launch pair."""
    from . import svg as psvg

    s = psvg.parseSvg(data, width, height)
    img = newImage(s.width, s.height)
    # the elements stay path COMMANDS: flattening, stroking and shapesToSegments run on the device (arcs, round caps /
    # joins and dashes are flattened by libpixie_host.so per path, see include/pixie_cuda.h)
    dev.CmdList.from_paths(s.width, s.height, 1, psvg.svg_path_batch(s)).run(img._d)
    return img
