"""Host-side SVG reader for the path-only subset the tiger needs.

Mirrors treeform/pixie src/pixie/fileformats/svg.nim for <svg>, <g>, <path> with the presentation
attributes fill, stroke, stroke-width, stroke-linecap, stroke-linejoin, stroke-miterlimit,
fill-rule, opacity, fill-opacity, stroke-opacity and transform=matrix()/translate()/scale()
(parseSvgProperties :54-300, parseSvg :502-555) and the render loop of newImage(svg) :557-608:
fillPath (first element OverwriteBlend, then NormalBlend) followed by strokePath.  It stays on the
host in the reference as well; this module only produces the ordered fill command list that the
C ABI (pixie_cuda_fill_batch) consumes.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field, replace

import numpy as np

from . import host
from .common import NormalBlend, OverwriteBlend, PixieError, parseHtmlColor, rgba_to_rgbx
from .device import FillBatch, PathBatch


@dataclass
class SvgProperties:  # svg.nim:19-30, initSvgProperties :44-52
    display: bool = True
    fillRule: int = host.NonZero
    fill: str = "black"
    stroke: int = 0  # premultiplied rgbx, 0 = none
    strokeWidth: float = 1.0
    strokeLineCap: int = host.ButtCap
    strokeLineJoin: int = host.MiterJoin
    strokeMiterLimit: float = host.defaultMiterLimit
    strokeDashArray: tuple = ()
    transform: np.ndarray = field(default_factory=host.mat3)
    opacity: float = 1.0
    fillOpacity: float = 1.0
    strokeOpacity: float = 1.0


def _tag(node):
    return node.tag.split("}")[-1]


def _split_args(s):
    return [e for e in s.replace(",", " ").split(" ") if e]


def _parse_props(node, inherited: SvgProperties) -> SvgProperties:
    r = replace(inherited)
    get = lambda k: node.attrib.get(k, "")
    attrs = {k: get(k) for k in ("fill-rule", "fill", "stroke", "stroke-width", "stroke-linecap", "stroke-linejoin",
                                 "stroke-miterlimit", "stroke-dasharray", "transform", "display", "opacity",
                                 "fill-opacity", "stroke-opacity")}
    for pair in get("style").split(";"):  # element attributes win over style (svg.nim:97-138)
        parts = pair.split(":")
        if len(parts) == 2:
            k = parts[0].strip()
            if k in attrs and k != "transform" and attrs[k] == "":
                attrs[k] = parts[1].strip()
    if attrs["display"]:
        r.display = attrs["display"].strip() != "none"
    if attrs["opacity"]:
        r.opacity = min(max(float(attrs["opacity"]), 0.0), 1.0)
    fr = attrs["fill-rule"]
    if fr == "nonzero":
        r.fillRule = host.NonZero
    elif fr == "evenodd":
        r.fillRule = host.EvenOdd
    elif fr != "":
        raise PixieError("Invalid fill-rule value " + fr)
    if attrs["fill"] not in ("", "currentColor"):
        r.fill = attrs["fill"]
    st = attrs["stroke"]
    if st == "currentColor":
        if r.stroke == 0:
            r.stroke = rgba_to_rgbx(0, 0, 0, 255)
    elif st == "none":
        r.stroke = 0
    elif st != "":
        r.stroke = rgba_to_rgbx(*parseHtmlColor(st))
    if attrs["fill-opacity"]:
        r.fillOpacity = min(max(float(attrs["fill-opacity"]), 0.0), 1.0)
    if attrs["stroke-opacity"]:
        r.strokeOpacity = min(max(float(attrs["stroke-opacity"]), 0.0), 1.0)
    sw = attrs["stroke-width"]
    if sw != "":
        if sw.endswith("px"):
            sw = sw[:-2]
        r.strokeWidth = float(np.float32(float(sw)))
        if r.stroke == 0:
            r.stroke = rgba_to_rgbx(0, 0, 0, 255)
    caps = {"butt": host.ButtCap, "round": host.RoundCap, "square": host.SquareCap}
    joins = {"miter": host.MiterJoin, "round": host.RoundJoin, "bevel": host.BevelJoin}
    if attrs["stroke-linecap"] not in ("", "inherit"):
        if attrs["stroke-linecap"] not in caps:
            raise PixieError("Invalid stroke-linecap value " + attrs["stroke-linecap"])
        r.strokeLineCap = caps[attrs["stroke-linecap"]]
    if attrs["stroke-linejoin"] not in ("", "inherit"):
        if attrs["stroke-linejoin"] not in joins:
            raise PixieError("Invalid stroke-linejoin value " + attrs["stroke-linejoin"])
        r.strokeLineJoin = joins[attrs["stroke-linejoin"]]
    if attrs["stroke-miterlimit"]:
        r.strokeMiterLimit = float(attrs["stroke-miterlimit"])
    if attrs["stroke-dasharray"]:
        r.strokeDashArray = tuple(r.strokeDashArray) + tuple(float(v) for v in _split_args(attrs["stroke-dasharray"]))
    tr = attrs["transform"]
    remaining = tr
    while remaining:
        idx = remaining.find(")")
        if idx == -1:
            raise PixieError("Unsupported SVG transform: " + tr)
        f = remaining[:idx + 1].strip()
        remaining = remaining[idx + 1:]
        if f.startswith("matrix("):
            arr = _split_args(f[7:-1])
            if len(arr) != 6:
                raise PixieError("Unsupported SVG transform: " + tr)
            m = host.mat3()
            m[0], m[1], m[3], m[4], m[6], m[7] = [np.float32(float(a)) for a in arr]
            r.transform = host.matmul(r.transform, m)
        elif f.startswith("translate("):
            c = _split_args(f[10:-1])
            r.transform = host.matmul(r.transform, host.translate(float(c[0]), float(c[1]) if len(c) > 1 else 0.0))
        elif f.startswith("scale("):
            c = _split_args(f[6:-1])
            sx = float(c[0])
            r.transform = host.matmul(r.transform, host.scale(sx, float(c[1]) if len(c) > 1 else sx))
        else:
            raise PixieError("Unsupported SVG transform: " + tr)
    return r


@dataclass
class Svg:
    width: int
    height: int
    elements: list  # [(path string, SvgProperties)]


def parseSvg(data: str, width: int = 0, height: int = 0) -> Svg:
    root = ET.fromstring(data)
    if _tag(root) != "svg":
        raise PixieError("Invalid SVG data")
    box = root.attrib.get("viewBox", "").split(" ")
    vbx, vby, vbw, vbh = (int(v) for v in box)
    props = _parse_props(root, SvgProperties())
    if vbx != 0 or vby != 0:
        props.transform = host.matmul(props.transform, host.translate(-float(vbx), -float(vby)))
    if width == 0 and height == 0:
        width, height = vbw, vbh
    else:
        sx = np.float32(width) / np.float32(vbw)
        sy = np.float32(height) / np.float32(vbh)
        props.transform = host.matmul(props.transform, host.scale(sx, sy))
    elements = []

    def walk(node, stack):
        t = _tag(node)
        if t in ("title", "desc", "defs"):
            return
        if t == "g":
            p = _parse_props(node, stack[-1])
            stack.append(p)
            for ch in node:
                walk(ch, stack)
            stack.pop()
        elif t == "path":
            elements.append((node.attrib.get("d", ""), _parse_props(node, stack[-1])))
        else:
            raise PixieError("Unsupported SVG tag: " + t + " (only the path subset is mirrored)")

    stack = [props]
    for ch in root:
        walk(ch, stack)
    return Svg(width, height, elements)


def _scaled_alpha(rgbx: int, factor: float) -> int:
    """paint from a ColorRGBX; color.a *= factor; asRgbx() (svg.nim:583-594, paths.nim:2110-2112)."""
    a = (rgbx >> 24) & 255
    if a == 0:
        return 0
    f32 = np.float32
    # rgbx -> Color (un-premultiply), scale alpha, -> rgbx.  For opaque colours and factor 1 this is the identity.
    if a == 255 and factor == 1.0:
        return rgbx
    ch = [f32((rgbx >> s) & 255) / f32(255) / (f32(a) / f32(255)) for s in (0, 8, 16)]
    na = f32(a) / f32(255) * f32(factor)
    q = lambda v: int(np.floor(float(f32(v) * f32(255)) + 0.5))
    r8, g8, b8, a8 = q(ch[0]), q(ch[1]), q(ch[2]), q(na)
    return rgba_to_rgbx(min(r8, 255), min(g8, 255), min(b8, 255), a8)


def svg_fill_batch(svg: Svg, layer: int = 0, batch: FillBatch | None = None) -> FillBatch:
    """The render loop of newImage(svg) (svg.nim:557-608) as an ordered command list."""
    b = batch if batch is not None else FillBatch()
    blend = OverwriteBlend
    for d, props in svg.elements:
        if not (props.display and props.opacity > 0):
            continue
        path = host.parsePath(d)
        if props.fill != "none":
            if props.fill.startswith("url("):
                raise PixieError("gradient fills are not on this path")
            opacity = max(0.0, min(1.0, props.fillOpacity * props.opacity))
            if opacity != 0:
                rgbx = _scaled_alpha(rgba_to_rgbx(*parseHtmlColor(props.fill)), opacity)
                if (rgbx >> 24) > 0 or blend == OverwriteBlend:
                    b.add(host.fill_segments(path, props.transform), rgbx, props.fillRule, blend, layer)
        blend = NormalBlend
        if props.stroke != 0 and props.strokeWidth > 0:
            rgbx = _scaled_alpha(props.stroke, props.opacity * props.strokeOpacity)
            if (rgbx >> 24) > 0:
                b.add(host.stroke_segments(path, props.transform, props.strokeWidth, props.strokeLineCap,
                                           props.strokeLineJoin, props.strokeMiterLimit, props.strokeDashArray),
                      rgbx, host.NonZero, NormalBlend, layer)
    return b


def svg_path_batch(svg: Svg, layer: int = 0, batch: PathBatch | None = None) -> PathBatch:
    """The same render loop with the paths left as COMMANDS: flattening, stroking and shapesToSegments run on the device
    (PathBatch -> pixie_cuda_cmdlist_create_from_paths)."""
    b = batch if batch is not None else PathBatch()
    blend = OverwriteBlend
    for d, props in svg.elements:
        if not (props.display and props.opacity > 0):
            continue
        path = host.parsePath(d)
        if props.fill != "none":
            if props.fill.startswith("url("):
                raise PixieError("gradient fills are not on this path")
            opacity = max(0.0, min(1.0, props.fillOpacity * props.opacity))
            if opacity != 0:
                rgbx = _scaled_alpha(rgba_to_rgbx(*parseHtmlColor(props.fill)), opacity)
                if (rgbx >> 24) > 0 or blend == OverwriteBlend:
                    b.add_fill(path, props.transform, rgbx, props.fillRule, blend, layer)
        blend = NormalBlend
        if props.stroke != 0 and props.strokeWidth > 0:
            rgbx = _scaled_alpha(props.stroke, props.opacity * props.strokeOpacity)
            if (rgbx >> 24) > 0:
                b.add_stroke(path, props.transform, props.strokeWidth, props.strokeLineCap, props.strokeLineJoin,
                             props.strokeMiterLimit, props.strokeDashArray, rgbx, host.NonZero, NormalBlend, layer)
    return b
