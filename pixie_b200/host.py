"""ctypes binding of libpixie_host.so — the C++ mirror of Pixie's Nim host-side producers.

Mirrors (names and argument meaning) treeform/pixie src/pixie/paths.nim: ``parsePath`` :119,
``Path`` builders :339-652, and the producer chains of ``fillPath`` :2108-2109 /
``strokePath`` :2163-2172 down to ``shapesToSegments`` :1059-1090; plus
``gaussianKernel`` (src/pixie/internal.nim:17-34).  CPU only; no GPU is touched here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .common import PixieError

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpixie_host.so")

ButtCap, RoundCap, SquareCap = 0, 1, 2          # paths.nim:10-12
MiterJoin, RoundJoin, BevelJoin = 0, 1, 2       # paths.nim:14-16
NonZero, EvenOdd = 0, 1                         # paths.nim:5-8
defaultMiterLimit = 4.0                         # paths.nim:46

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_LIB_PATH)
        f32, i32, vp = C.c_float, C.c_int, C.c_void_p
        L.pixie_host_last_error.restype = C.c_char_p
        L.pixie_host_path_new.restype = vp
        L.pixie_host_path_free.argtypes = [vp]
        L.pixie_host_path_parse.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.pixie_host_path_num_commands.argtypes = [vp]
        L.pixie_host_path_commands.argtypes = [vp, vp, i32]
        L.pixie_host_path_move_to.argtypes = [vp, f32, f32]
        L.pixie_host_path_line_to.argtypes = [vp, f32, f32]
        L.pixie_host_path_bezier_curve_to.argtypes = [vp] + [f32] * 6
        L.pixie_host_path_quadratic_curve_to.argtypes = [vp] + [f32] * 4
        L.pixie_host_path_elliptical_arc_to.argtypes = [vp, f32, f32, f32, i32, i32, f32, f32]
        L.pixie_host_path_arc.argtypes = [vp, f32, f32, f32, f32, f32, i32]
        L.pixie_host_path_arc_to.argtypes = [vp] + [f32] * 5
        L.pixie_host_path_rect.argtypes = [vp, f32, f32, f32, f32, i32]
        L.pixie_host_path_rounded_rect.argtypes = [vp] + [f32] * 8 + [i32]
        L.pixie_host_path_ellipse.argtypes = [vp] + [f32] * 4
        L.pixie_host_path_polygon.argtypes = [vp, f32, f32, f32, i32]
        L.pixie_host_path_close.argtypes = [vp]
        L.pixie_host_fill_segments.argtypes = [vp, vp, C.POINTER(vp)]
        L.pixie_host_stroke_segments.argtypes = [vp, vp, f32, i32, i32, f32, vp, i32, C.POINTER(vp)]
        L.pixie_host_segments_count.argtypes = [vp]
        L.pixie_host_segments_xyxy.argtypes = [vp]
        L.pixie_host_segments_xyxy.restype = vp
        L.pixie_host_segments_winding.argtypes = [vp]
        L.pixie_host_segments_winding.restype = vp
        L.pixie_host_segments_free.argtypes = [vp]
        L.pixie_host_gaussian_kernel.argtypes = [i32, vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PixieError(lib().pixie_host_last_error().decode())


# ----------------------------------------------------------------------------- vmath Mat3
def mat3():
    return np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], dtype=np.float32)


def translate(x, y):
    m = mat3()
    m[6], m[7] = x, y
    return m


def scale(x, y):
    m = mat3()
    m[0], m[4] = x, y
    return m


def rotate(angle):
    s, c = np.float32(np.sin(np.float32(angle))), np.float32(np.cos(np.float32(angle)))
    return np.array([c, s, 0, -s, c, 0, 0, 0, 1], dtype=np.float32)


def matmul(a, b):
    """vmath `*`(a, b: Mat3): result column i = a * (column i of b), float32, left-to-right sums."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    r = np.zeros(9, dtype=np.float32)
    for c in range(3):
        for row in range(3):
            acc = np.float32(b[c * 3 + 0] * a[0 * 3 + row])
            acc = np.float32(acc + np.float32(b[c * 3 + 1] * a[1 * 3 + row]))
            acc = np.float32(acc + np.float32(b[c * 3 + 2] * a[2 * 3 + row]))
            r[c * 3 + row] = acc
    return r


class Segments:
    """seq[(Segment, int16)] as two arrays: xyxy float32 [n,4] (at.y < to.y), winding int16 [n]."""

    __slots__ = ("xyxy", "winding")

    def __init__(self, xyxy, winding):
        self.xyxy = np.ascontiguousarray(xyxy, dtype=np.float32).reshape(-1, 4)
        self.winding = np.ascontiguousarray(winding, dtype=np.int16).reshape(-1)

    def __len__(self):
        return int(self.winding.shape[0])


def _take_segments(handle):
    L = lib()
    n = L.pixie_host_segments_count(handle)
    if n:
        xy = np.ctypeslib.as_array(C.cast(L.pixie_host_segments_xyxy(handle), C.POINTER(C.c_float)), (n, 4)).copy()
        w = np.ctypeslib.as_array(C.cast(L.pixie_host_segments_winding(handle), C.POINTER(C.c_int16)), (n,)).copy()
    else:
        xy = np.zeros((0, 4), np.float32)
        w = np.zeros((0,), np.int16)
    L.pixie_host_segments_free(handle)
    return Segments(xy, w)


class Path:
    """paths.nim:24-27 ``Path`` (a float32 command stream)."""

    def __init__(self, _handle=None):
        self._h = _handle if _handle is not None else lib().pixie_host_path_new()

    def __del__(self):
        try:
            if self._h:
                lib().pixie_host_path_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def commands(self):
        n = lib().pixie_host_path_num_commands(self._h)
        out = np.zeros(n, np.float32)
        if n:
            lib().pixie_host_path_commands(self._h, out.ctypes.data, n)
        return out

    def moveTo(self, x, y): lib().pixie_host_path_move_to(self._h, x, y)
    def lineTo(self, x, y): lib().pixie_host_path_line_to(self._h, x, y)
    def bezierCurveTo(self, x1, y1, x2, y2, x3, y3): lib().pixie_host_path_bezier_curve_to(self._h, x1, y1, x2, y2, x3, y3)
    def quadraticCurveTo(self, x1, y1, x2, y2): lib().pixie_host_path_quadratic_curve_to(self._h, x1, y1, x2, y2)

    def ellipticalArcTo(self, rx, ry, xAxisRotation, largeArcFlag, sweepFlag, x, y):
        lib().pixie_host_path_elliptical_arc_to(self._h, rx, ry, xAxisRotation, int(largeArcFlag), int(sweepFlag), x, y)

    def arc(self, x, y, r, a0, a1, ccw=False): _check(lib().pixie_host_path_arc(self._h, x, y, r, a0, a1, int(ccw)))
    def arcTo(self, x1, y1, x2, y2, r): _check(lib().pixie_host_path_arc_to(self._h, x1, y1, x2, y2, r))
    def rect(self, x, y, w, h, clockwise=True): lib().pixie_host_path_rect(self._h, x, y, w, h, int(clockwise))

    def roundedRect(self, x, y, w, h, nw, ne, se, sw, clockwise=True):
        lib().pixie_host_path_rounded_rect(self._h, x, y, w, h, nw, ne, se, sw, int(clockwise))

    def ellipse(self, cx, cy, rx, ry): lib().pixie_host_path_ellipse(self._h, cx, cy, rx, ry)
    def circle(self, cx, cy, r): lib().pixie_host_path_ellipse(self._h, cx, cy, r, r)
    def polygon(self, x, y, size, sides): _check(lib().pixie_host_path_polygon(self._h, x, y, size, sides))
    def closePath(self): lib().pixie_host_path_close(self._h)


def newPath():
    return Path()


def parsePath(path: str) -> Path:
    h = C.c_void_p()
    _check(lib().pixie_host_path_parse(path.encode(), C.byref(h)))
    return Path(h)


def _some_path(path):
    return parsePath(path) if isinstance(path, str) else path


def _mat_ptr(transform):
    if transform is None:
        return None, None
    m = np.ascontiguousarray(transform, dtype=np.float32).reshape(9)
    return m, m.ctypes.data


def fill_segments(path, transform=None) -> Segments:
    """parseSomePath(closeSubpaths=true) -> transform -> shapesToSegments (paths.nim:2108-2109,1604)."""
    p = _some_path(path)
    keep, ptr = _mat_ptr(transform)
    h = C.c_void_p()
    _check(lib().pixie_host_fill_segments(p._h, ptr, C.byref(h)))
    return _take_segments(h)


def stroke_segments(path, transform=None, strokeWidth=1.0, lineCap=ButtCap, lineJoin=MiterJoin,
                    miterLimit=defaultMiterLimit, dashes=()) -> Segments:
    """strokeShapes(...) -> transform -> shapesToSegments (paths.nim:2163-2172,1604)."""
    p = _some_path(path)
    keep, ptr = _mat_ptr(transform)
    d = np.ascontiguousarray(list(dashes), dtype=np.float32)
    h = C.c_void_p()
    _check(lib().pixie_host_stroke_segments(p._h, ptr, strokeWidth, lineCap, lineJoin, miterLimit,
                                            d.ctypes.data if len(d) else None, len(d), C.byref(h)))
    return _take_segments(h)


def gaussianKernel(radius: int) -> np.ndarray:
    """internal.nim:17-34 — uint16 LUT with 2*radius+1 taps."""
    out = np.zeros(2 * radius + 1, np.uint16)
    _check(lib().pixie_host_gaussian_kernel(radius, out.ctypes.data))
    return out
