"""ctypes binding of the CPU oracle (oracle/libpixie_oracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libpixie_oracle.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(LIB)
        vp, i32, u32, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_float
        L.orc_last_error.restype = C.c_char_p
        L.orc_fill_segments.argtypes = [vp, i32, i32, vp, vp, i32, u32, i32, i32, i32, vp]
        L.orc_blend_px.argtypes = [i32, u32, u32]
        L.orc_blend_px.restype = u32
        L.orc_blend_rect.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, i32]
        L.orc_blend_rect_masked.argtypes = [vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32]
        L.orc_apply_opacity.argtypes = [vp, i32, i32, f32]
        L.orc_blur.argtypes = [vp, i32, i32, vp, i32, u32]
        L.orc_spread.argtypes = [vp, i32, i32, i32]
        L.orc_shadow.argtypes = [vp, i32, i32, f32, f32, i32, vp, i32, u32, vp]
        L.orc_minify_by2.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        L.orc_magnify_by2.argtypes = [vp, i32, i32, i32, vp]
        L.orc_draw.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32]
        L.orc_draw_correct.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32, i32]
        L.orc_fill_gradient.argtypes = [vp, i32, i32, i32, vp, i32, vp, vp, i32, f32]
        _lib = L
    return _lib


class OracleError(Exception):
    pass


def _chk(rc):
    if rc:
        raise OracleError(lib().orc_last_error().decode())


class OracleBackend:
    """Backend protocol of tests/golden_cases.py on the CPU oracle.  sem=0 canonical, 1 scalar."""

    def __init__(self, sem=0):
        self.sem = sem
        self.covered = 0

    def fill_segments(self, img, segs, rgbx, rule, mode):
        assert img.dtype == np.uint8 and img.flags.c_contiguous
        h, w = img.shape[:2]
        cov = C.c_uint64(0)
        _chk(lib().orc_fill_segments(img.ctypes.data, w, h, segs.xyxy.ctypes.data, segs.winding.ctypes.data,
                                     len(segs), rgbx, rule, mode, self.sem, C.addressof(cov)))
        self.covered += cov.value

    def blend_rect(self, dst, src, px, py, mode):
        _chk(lib().orc_blend_rect(dst.ctypes.data, dst.shape[1], dst.shape[0], src.ctypes.data, src.shape[1],
                                  src.shape[0], px, py, mode))

    def blend_rect_masked(self, dst, src, mask, px, py, mode):
        is_rgbx = 1 if mask.ndim == 3 else 0
        _chk(lib().orc_blend_rect_masked(dst.ctypes.data, dst.shape[1], dst.shape[0], src.ctypes.data,
                                         mask.ctypes.data, is_rgbx, src.shape[1], src.shape[0], px, py, mode))

    def blur(self, img, lut, radius, oob):
        lut = np.ascontiguousarray(lut, np.uint16)
        _chk(lib().orc_blur(img.ctypes.data, img.shape[1], img.shape[0], lut.ctypes.data, radius, oob))

    def spread(self, img, spread):
        _chk(lib().orc_spread(img.ctypes.data, img.shape[1], img.shape[0], spread))

    def shadow(self, img, ox, oy, spread, lut, radius, rgbx):
        lut = np.ascontiguousarray(lut, np.uint16)
        out = np.zeros_like(img)
        _chk(lib().orc_shadow(img.ctypes.data, img.shape[1], img.shape[0], ox, oy, spread, lut.ctypes.data, radius,
                              rgbx, out.ctypes.data))
        return out


    def apply_opacity(self, img, opacity):
        _chk(lib().orc_apply_opacity(img.ctypes.data, img.shape[1], img.shape[0], opacity))

    def draw(self, dst, src, mat, mode):
        m = np.ascontiguousarray(mat, np.float32)
        src = np.ascontiguousarray(src)
        _chk(lib().orc_draw(dst.ctypes.data, dst.shape[1], dst.shape[0], src.ctypes.data, src.shape[1], src.shape[0],
                            m.ctypes.data, mode))

    def draw_tiled(self, dst, src, mat, mode, tiled=True):
        m = np.ascontiguousarray(mat, np.float32)
        src = np.ascontiguousarray(src)
        _chk(lib().orc_draw_correct(dst.ctypes.data, dst.shape[1], dst.shape[0], src.ctypes.data, src.shape[1],
                                    src.shape[0], m.ctypes.data, mode, 1 if tiled else 0))

    def minify_by2(self, img, power=1):
        img = np.ascontiguousarray(img)
        ow, oh = C.c_int(0), C.c_int(0)
        _chk(lib().orc_minify_by2(img.ctypes.data, img.shape[1], img.shape[0], power, None, C.addressof(ow), C.addressof(oh)))
        out = np.zeros((oh.value, ow.value, 4), np.uint8)
        _chk(lib().orc_minify_by2(img.ctypes.data, img.shape[1], img.shape[0], power, out.ctypes.data, C.addressof(ow),
                                  C.addressof(oh)))
        return out

    def magnify_by2(self, img, power=1):
        img = np.ascontiguousarray(img)
        out = np.zeros((img.shape[0] << power, img.shape[1] << power, 4), np.uint8)
        _chk(lib().orc_magnify_by2(img.ctypes.data, img.shape[1], img.shape[0], power, out.ctypes.data))
        return out

    def fill_gradient(self, img, kind, handles, stops, opacity=1.0):
        """stops: [(position, (r, g, b, a))] with float colours (chroma Color)."""
        hx = np.ascontiguousarray(np.asarray(handles, np.float32).reshape(-1))
        pos = np.ascontiguousarray([s[0] for s in stops], np.float32)
        col = np.ascontiguousarray([s[1] for s in stops], np.float32).reshape(-1)
        _chk(lib().orc_fill_gradient(img.ctypes.data, img.shape[1], img.shape[0], kind, hx.ctypes.data, len(hx) // 2,
                                     pos.ctypes.data, col.ctypes.data, len(stops), opacity))


def blend_px(mode, backdrop, source):
    return lib().orc_blend_px(mode, backdrop, source)
