"""ctypes binding of pixie_cuda.so — the C ABI declared in include/pixie_cuda.h.

There is no CPU fallback: loading fails loudly when the CUDA library is missing, and every call
returns the library's status, raised here as PixieError (the analogue of the Nim shim's
``raise newException(PixieError, $pixie_cuda_last_error())``).
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

from .common import PixieError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIXIE_CUDA_LIB") or os.path.join(_HERE, "pixie_cuda.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pixie_cuda.h")

_lib = None

vp, i32, u32, u64, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
P = C.POINTER

_SIGNATURES = {
    "pixie_cuda_init": [i32],
    "pixie_cuda_set_stream": [vp],
    "pixie_cuda_sync": [],
    "pixie_cuda_set_sm_reserve": [i32],
    "pixie_cuda_device_count": [P(i32)],
    "pixie_cuda_image_create": [i32, i32, P(u64)],
    "pixie_cuda_image_create_layers": [i32, i32, i32, P(u64)],
    "pixie_cuda_image_create_a8": [i32, i32, P(u64)],
    "pixie_cuda_image_wrap": [vp, i32, i32, i32, i32, P(u64)],
    "pixie_cuda_image_destroy": [u64],
    "pixie_cuda_image_info": [u64, P(i32), P(i32), P(i32), P(i32), P(vp)],
    "pixie_cuda_image_upload": [u64, vp],
    "pixie_cuda_image_download": [u64, vp],
    "pixie_cuda_image_upload_async": [u64, vp],
    "pixie_cuda_image_download_async": [u64, vp],
    "pixie_cuda_image_download_rows": [u64, i32, i32, i32, vp],
    "pixie_cuda_image_fill": [u64, u32],
    "pixie_cuda_image_copy": [u64, u64],
    "pixie_cuda_image_checksum": [u64, P(u64)],
    "pixie_cuda_fill_segments": [u64, vp, vp, i32, u32, i32, i32],
    "pixie_cuda_fill_batch": [u64, i32, vp, vp, vp, vp, vp, vp, vp, P(u64)],
    "pixie_cuda_render_batch_host": [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, P(u64)],
    "pixie_cuda_cmdlist_create": [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, P(u64)],
    "pixie_cuda_cmdlist_run": [u64, u64, P(u64)],
    "pixie_cuda_cmdlist_run_cleared": [u64, u64, P(u64)],
    "pixie_cuda_cmdlist_run_rows": [u64, u64, i32, i32, P(u64)],
    "pixie_cuda_cmdlist_info": [u64, P(C.c_int64), P(C.c_int64), P(C.c_int64), P(C.c_int64)],
    "pixie_cuda_cmdlist_destroy": [u64],
    "pixie_cuda_cmdlist_create_from_paths": [i32, i32, i32, i32, vp, vp, C.c_int64, vp, vp, C.c_int64, P(u64)],
    "pixie_cuda_cmdlist_segments": [u64, vp, vp, vp],
    "pixie_cuda_render_paths_host": [vp, i32, i32, i32, i32, vp, vp, C.c_int64, vp, vp, C.c_int64, P(u64)],
    "pixie_cuda_cmdlist_set_overlap": [u64, i32],
    "pixie_cuda_blend_rect": [u64, u64, i32, i32, i32],
    "pixie_cuda_blend_rect_masked": [u64, u64, u64, i32, i32, i32],
    "pixie_cuda_apply_opacity": [u64, f32],
    "pixie_cuda_draw": [u64, u64, vp, i32],
    "pixie_cuda_draw_tiled": [u64, u64, vp, i32],
    "pixie_cuda_draw_correct": [u64, u64, vp, i32],
    "pixie_cuda_minify_by2": [u64, i32, P(u64)],
    "pixie_cuda_magnify_by2": [u64, i32, P(u64)],
    "pixie_cuda_fill_gradient": [u64, i32, vp, i32, vp, vp, i32, f32],
    "pixie_cuda_fill_gradient_masked": [u64, u64, i32, vp, i32, vp, vp, i32, f32, i32],
    "pixie_cuda_blur": [u64, vp, i32, u32],
    "pixie_cuda_blur_rows": [u64, vp, i32, u32, i32, i32],
    "pixie_cuda_blur_rows_to": [u64, u64, vp, i32, u32, i32, i32],
    "pixie_cuda_blur_rows_to_flags": [u64, u64, vp, i32, u32, i32, i32, vp, vp, u32],
    "pixie_cuda_blur_rows_x": [u64, vp, i32, u32, i32, i32],
    "pixie_cuda_blur_rows_y": [u64, vp, i32, u32, i32, i32],
    "pixie_cuda_spread": [u64, i32],
    "pixie_cuda_spread_rows": [u64, i32, i32, i32],
    "pixie_cuda_shadow_rows": [u64, u64, f32, f32, i32, vp, i32, u32, i32, i32],
    "pixie_cuda_spread_host": [vp, i32, i32, i32],
    "pixie_cuda_apply_opacity_host": [vp, i32, i32, f32],
    "pixie_cuda_blend_rect_masked_host": [vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32],
    "pixie_cuda_shadow": [u64, u64, f32, f32, i32, vp, i32, u32],
    "pixie_cuda_fill_segments_host": [vp, i32, i32, vp, vp, i32, u32, i32, i32],
    "pixie_cuda_blend_rect_host": [vp, i32, i32, vp, i32, i32, i32, i32, i32],
    "pixie_cuda_blur_host": [vp, i32, i32, vp, i32, u32],
    "pixie_cuda_shadow_host": [vp, vp, i32, i32, f32, f32, i32, vp, i32, u32],
    "pixie_cuda_draw_host": [vp, i32, i32, vp, i32, i32, vp, i32, i32],
    "pixie_cuda_fill_gradient_host": [vp, i32, i32, i32, vp, i32, vp, vp, i32, f32],
    "pixie_cuda_minify_by2_host": [vp, i32, i32, i32, vp],
    "pixie_cuda_magnify_by2_host": [vp, i32, i32, i32, vp],
    "pixie_cuda_peer_alloc": [C.c_size_t, P(vp), vp],
    "pixie_cuda_peer_open": [vp, P(vp)],
    "pixie_cuda_peer_close": [vp],
    "pixie_cuda_peer_free": [vp],
    "pixie_cuda_halo_push": [vp, vp, C.c_size_t, vp, u32],
    "pixie_cuda_halo_wait": [vp, u32],
    "pixie_cuda_halo_exchange": [vp, vp, u32],
    "pixie_cuda_halo_wait2": [vp, vp, u32],
    "pixie_cuda_host_alloc": [C.c_size_t, P(vp)],
    "pixie_cuda_host_free": [vp],
    "pixie_cuda_set_profiling": [i32],
    "pixie_cuda_profile_read": [i32, P(f32)],
    "pixie_cuda_launch_count": [P(u64)],
    "pixie_cuda_timer_begin": [],
    "pixie_cuda_timer_end": [P(f32)],
}


def declared_symbols():
    """Every function name include/pixie_cuda.h declares."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pixie_cuda_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load pixie_cuda.so (fails loudly if absent) and bind every symbol the header declares."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "pixie_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name in declared_symbols():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            if name == "pixie_cuda_last_error":
                fn.restype = C.c_char_p
                fn.argtypes = []
            else:
                fn.restype = i32
                fn.argtypes = _SIGNATURES[name]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PixieError(lib().pixie_cuda_last_error().decode())


_device = None


def init(device: int | None = None):
    """Select the GPU of this process (one process per GPU).  `None` = the device already initialised, else 0.
    Initialising a second, different device is an error (pixie_cuda_init)."""
    global _device
    if device is None:
        device = _device if _device is not None else 0
    check(lib().pixie_cuda_init(device))
    _device = device


def current_device():
    return _device


def sync():
    check(lib().pixie_cuda_sync())


def set_sm_reserve(sms: int):
    check(lib().pixie_cuda_set_sm_reserve(sms))


def set_stream(cuda_stream_ptr):
    check(lib().pixie_cuda_set_stream(cuda_stream_ptr))


def launch_count() -> int:
    v = u64(0)
    check(lib().pixie_cuda_launch_count(C.byref(v)))
    return v.value


def device_count() -> int:
    v = i32(0)
    rc = lib().pixie_cuda_device_count(C.byref(v))
    return v.value if rc == 0 else 0


def _ptr(a):
    return a.ctypes.data if a is not None else None


class DeviceImage:
    """A device-resident canvas (RGBX, optionally several independent layers) or A8 coverage plane."""

    def __init__(self, width, height, layers=1, a8=False, _handle=None, _owner=None):
        self.width, self.height, self.layers, self.a8 = width, height, layers, a8
        self._owner = _owner
        if _handle is not None:
            self.handle = _handle
            return
        h = u64(0)
        if a8:
            check(lib().pixie_cuda_image_create_a8(width, height, C.byref(h)))
        elif layers == 1:
            check(lib().pixie_cuda_image_create(width, height, C.byref(h)))
        else:
            check(lib().pixie_cuda_image_create_layers(width, height, layers, C.byref(h)))
        self.handle = h.value

    @classmethod
    def wrap(cls, device_ptr, width, height, layers=1, bytes_per_pixel=4, owner=None):
        h = u64(0)
        check(lib().pixie_cuda_image_wrap(device_ptr, width, height, layers, bytes_per_pixel, C.byref(h)))
        return cls(width, height, layers, bytes_per_pixel == 1, _handle=h.value, _owner=owner)

    def __del__(self):
        try:
            if getattr(self, "handle", 0):
                lib().pixie_cuda_image_destroy(self.handle)
                self.handle = 0
        except Exception:
            pass

    @property
    def shape(self):
        base = (self.height, self.width) if self.a8 else (self.height, self.width, 4)
        return base if self.layers == 1 else (self.layers,) + base

    def upload(self, pixels: np.ndarray):
        a = np.ascontiguousarray(pixels, dtype=np.uint8)
        assert a.size == int(np.prod(self.shape)), (a.shape, self.shape)
        check(lib().pixie_cuda_image_upload(self.handle, a.ctypes.data))
        return self

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, np.uint8)
        check(lib().pixie_cuda_image_download(self.handle, out.ctypes.data))
        return out

    def download_rows(self, y0, y1, layer=0):
        out = np.empty((y1 - y0, self.width) + (() if self.a8 else (4,)), np.uint8)
        check(lib().pixie_cuda_image_download_rows(self.handle, layer, y0, y1, out.ctypes.data))
        return out

    def fill(self, rgbx: int):
        check(lib().pixie_cuda_image_fill(self.handle, rgbx))

    def copy_from(self, other: "DeviceImage"):
        check(lib().pixie_cuda_image_copy(self.handle, other.handle))

    def checksum(self) -> int:
        v = u64(0)
        check(lib().pixie_cuda_image_checksum(self.handle, C.byref(v)))
        return v.value

    def device_ptr(self) -> int:
        p = vp()
        check(lib().pixie_cuda_image_info(self.handle, None, None, None, None, C.byref(p)))
        return p.value


class FillBatch:
    """Host-side ordered fill command list in the layout pixie_cuda_fill_batch takes."""

    def __init__(self):
        self._segs, self._wind = [], []
        self.seg_offsets = [0]
        self.rgbx, self.rule, self.mode, self.layer = [], [], [], []

    def add(self, segs, rgbx, rule, mode, layer=0):
        self._segs.append(segs.xyxy)
        self._wind.append(segs.winding)
        self.seg_offsets.append(self.seg_offsets[-1] + len(segs))
        self.rgbx.append(rgbx)
        self.rule.append(rule)
        self.mode.append(mode)
        self.layer.append(layer)

    def __len__(self):
        return len(self.rgbx)

    def arrays(self):
        xy = np.ascontiguousarray(np.concatenate(self._segs, axis=0) if self._segs else np.zeros((0, 4)), np.float32)
        wd = np.ascontiguousarray(np.concatenate(self._wind) if self._wind else np.zeros((0,)), np.int16)
        return dict(
            xyxy=xy, winding=wd, seg_offsets=np.asarray(self.seg_offsets, np.int32),
            rgbx=np.asarray(self.rgbx, np.uint32), rule=np.asarray(self.rule, np.uint8),
            mode=np.asarray(self.mode, np.uint8), layer=np.asarray(self.layer, np.int32))


def fill_segments(image: DeviceImage, segs, rgbx, rule, mode):
    check(lib().pixie_cuda_fill_segments(image.handle, _ptr(segs.xyxy), _ptr(segs.winding), len(segs), rgbx, rule, mode))


def fill_batch(image: DeviceImage, arrays: dict, count_covered=False):
    cov = u64(0)
    check(lib().pixie_cuda_fill_batch(
        image.handle, len(arrays["rgbx"]), _ptr(arrays["layer"]), _ptr(arrays["xyxy"]), _ptr(arrays["winding"]),
        _ptr(arrays["seg_offsets"]), _ptr(arrays["rgbx"]), _ptr(arrays["rule"]), _ptr(arrays["mode"]),
        C.byref(cov) if count_covered else None))
    return cov.value


def render_batch_host(pixels_ptr: int, width: int, height: int, arrays: dict, clear=True, count_covered=False):
    """pixie_cuda_render_batch_host: fresh canvas + ordered fills + pixels back to host memory at `pixels_ptr`
    (ideally a PinnedBuffer).  Single-layer command lists only."""
    cov = u64(0)
    check(lib().pixie_cuda_render_batch_host(
        pixels_ptr, width, height, 1 if clear else 0, len(arrays["rgbx"]), _ptr(arrays["xyxy"]), _ptr(arrays["winding"]),
        _ptr(arrays["seg_offsets"]), _ptr(arrays["rgbx"]), _ptr(arrays["rule"]), _ptr(arrays["mode"]),
        C.byref(cov) if count_covered else None))
    return cov.value


class PathDesc(C.Structure):  # pixie_path_desc (include/pixie_cuda.h)
    _fields_ = [("kind", C.c_int32), ("begin", C.c_int32), ("end", C.c_int32), ("num_commands", C.c_int32),
                ("transform", C.c_float * 9), ("stroke_width", C.c_float), ("line_cap", C.c_int32), ("line_join", C.c_int32),
                ("miter_limit", C.c_float), ("rgbx", C.c_uint32), ("winding_rule", C.c_uint8), ("blend_mode", C.c_uint8),
                ("reserved", C.c_uint16), ("layer", C.c_int32)]


_PARAMS = (0, 2, 2, 1, 1, 6, 4, 4, 2, 7, 2, 2, 1, 1, 6, 4, 4, 2, 7)  # parameterCount per PathCommandKind, paths.nim:73-81
_ARC_KINDS = (9, 18)


def _scan_commands(cmds):
    """(number of commands, whether an arc is among them) of a Path.commands stream."""
    i, n, arc = 0, 0, False
    while i < len(cmds):
        k = int(cmds[i])
        if k < 0 or k >= len(_PARAMS):
            raise PixieError("Invalid path command")
        arc = arc or k in _ARC_KINDS
        i += 1 + _PARAMS[k]
        n += 1
    return n, arc


class PathBatch:
    """Ordered list of fillPath / strokePath calls as path COMMANDS: flattening, stroking and shapesToSegments run on
    the device (pixie_cuda_cmdlist_create_from_paths).  Paths with arcs, round caps / joins or dashes are flattened by
    libpixie_host.so here and passed through as finished segments (the header says why)."""

    def __init__(self):
        self.descs, self._cmds, self._raw, self._rawWind = [], [], [], []
        self.ncmd, self.nraw, self.host_paths = 0, 0, 0

    def __len__(self):
        return len(self.descs)

    def _desc(self, kind, begin, end, ncommands, transform, rgbx, rule, mode, layer, strokeWidth=1.0, cap=0, join=0, miter=4.0):
        d = PathDesc()
        d.kind, d.begin, d.end, d.num_commands = kind, begin, end, ncommands
        m = np.eye(3, dtype=np.float32).reshape(9) if transform is None else np.ascontiguousarray(transform, np.float32).reshape(9)
        d.transform = (C.c_float * 9)(*[float(v) for v in m])
        d.stroke_width, d.line_cap, d.line_join, d.miter_limit = float(strokeWidth), int(cap), int(join), float(miter)
        d.rgbx, d.winding_rule, d.blend_mode, d.layer = int(rgbx), int(rule), int(mode), int(layer)
        self.descs.append(d)

    def _add_raw(self, segs, rgbx, rule, mode, layer):
        self._raw.append(segs.xyxy)
        self._rawWind.append(segs.winding)
        self._desc(2, self.nraw, self.nraw + len(segs), 0, None, rgbx, rule, mode, layer)
        self.nraw += len(segs)
        self.host_paths += 1

    def add_fill(self, path, transform, rgbx, rule, mode, layer=0):
        from . import host

        cmds = path.commands
        n, arc = _scan_commands(cmds)
        if arc:
            return self._add_raw(host.fill_segments(path, transform), rgbx, rule, mode, layer)
        self._cmds.append(cmds)
        self._desc(0, self.ncmd, self.ncmd + len(cmds), n, transform, rgbx, rule, mode, layer)
        self.ncmd += len(cmds)

    def add_stroke(self, path, transform, strokeWidth, lineCap, lineJoin, miterLimit, dashes, rgbx, rule, mode, layer=0):
        from . import host

        cmds = path.commands
        n, arc = _scan_commands(cmds)
        if arc or lineCap == host.RoundCap or lineJoin == host.RoundJoin or len(dashes) or not strokeWidth > 0:
            segs = host.stroke_segments(path, transform, strokeWidth, lineCap, lineJoin, miterLimit, dashes)
            return self._add_raw(segs, rgbx, rule, mode, layer)
        self._cmds.append(cmds)
        self._desc(1, self.ncmd, self.ncmd + len(cmds), n, transform, rgbx, rule, mode, layer, strokeWidth, lineCap, lineJoin, miterLimit)
        self.ncmd += len(cmds)

    def packed(self):
        descs = (PathDesc * max(1, len(self.descs)))(*self.descs)
        cmds = np.ascontiguousarray(np.concatenate(self._cmds) if self._cmds else np.zeros(0), np.float32)
        raw = np.ascontiguousarray(np.concatenate(self._raw, axis=0) if self._raw else np.zeros((0, 4)), np.float32)
        rw = np.ascontiguousarray(np.concatenate(self._rawWind) if self._rawWind else np.zeros(0), np.int16)
        return descs, cmds, raw, rw


def render_paths_host(pixels_ptr: int, width: int, height: int, batch: "PathBatch", packed=None, clear=True, count_covered=False):
    """pixie_cuda_render_paths_host: path commands in, flattened / stroked / rasterised on the device, pixels back to
    host memory at `pixels_ptr` (ideally a PinnedBuffer)."""
    descs, cmds, raw, rw = packed if packed is not None else batch.packed()
    cov = u64(0)
    check(lib().pixie_cuda_render_paths_host(
        pixels_ptr, width, height, 1 if clear else 0, len(batch), C.cast(descs, vp), _ptr(cmds), len(cmds), _ptr(raw), _ptr(rw),
        len(rw), C.byref(cov) if count_covered else None))
    return cov.value


class CmdList:
    """Device-resident command list (segments + fill headers in HBM)."""

    @classmethod
    def from_paths(cls, width, height, layers, batch: "PathBatch", packed=None):
        """Path commands in, flattened / stroked on the device (pixie_cuda_cmdlist_create_from_paths)."""
        descs, cmds, raw, rw = packed if packed is not None else batch.packed()
        self = cls.__new__(cls)
        h = u64(0)
        check(lib().pixie_cuda_cmdlist_create_from_paths(
            width, height, layers, len(batch), C.cast(descs, vp), _ptr(cmds), len(cmds), _ptr(raw), _ptr(rw), len(rw), C.byref(h)))
        self.handle = h.value
        self.num_fills = len(batch)
        return self

    def segments(self, num_fills=None):
        """The list's segments back on the host: (xyxy [n, 4] float32, winding [n] int16, seg_offsets [fills + 1])."""
        n = self.info()["segments"]
        nf = self.num_fills if num_fills is None else num_fills
        xy, wd, so = np.zeros((n, 4), np.float32), np.zeros(n, np.int16), np.zeros(nf + 1, np.int32)
        check(lib().pixie_cuda_cmdlist_segments(self.handle, _ptr(xy), _ptr(wd), _ptr(so)))
        return xy, wd, so

    def __init__(self, width, height, layers, arrays: dict):
        self.num_fills = len(arrays["rgbx"])
        h = u64(0)
        check(lib().pixie_cuda_cmdlist_create(
            width, height, layers, len(arrays["rgbx"]), _ptr(arrays["layer"]), _ptr(arrays["xyxy"]),
            _ptr(arrays["winding"]), _ptr(arrays["seg_offsets"]), _ptr(arrays["rgbx"]), _ptr(arrays["rule"]),
            _ptr(arrays["mode"]), C.byref(h)))
        self.handle = h.value

    def run(self, image: DeviceImage, count_covered=False, clear=False):
        """clear=True: onto a transparent canvas (newImage + fills); the raster kernel clears the canvas as it goes."""
        cov = u64(0)
        fn = lib().pixie_cuda_cmdlist_run_cleared if clear else lib().pixie_cuda_cmdlist_run
        check(fn(self.handle, image.handle, C.byref(cov) if count_covered else None))
        return cov.value

    def run_rows(self, image: DeviceImage, y0, y1, count_covered=False):
        """Rows [y0, y1) only; `image` = the whole canvas or a band image of y1 - y0 rows (multi-GPU row bands)."""
        cov = u64(0)
        check(lib().pixie_cuda_cmdlist_run_rows(self.handle, image.handle, y0, y1, C.byref(cov) if count_covered else None))
        return cov.value

    def set_overlap(self, enabled: bool):
        """False: run a banded list as one band, kernels one after the other (to time a kernel alone)."""
        check(lib().pixie_cuda_cmdlist_set_overlap(self.handle, 1 if enabled else 0))

    def info(self):
        a, b, c, d = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(lib().pixie_cuda_cmdlist_info(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(segments=a.value, partitions=b.value, entries=c.value, launches_per_run=d.value)

    def __del__(self):
        try:
            if getattr(self, "handle", 0):
                lib().pixie_cuda_cmdlist_destroy(self.handle)
                self.handle = 0
        except Exception:
            pass


def blend_rect(dst: DeviceImage, src: DeviceImage, px, py, mode):
    check(lib().pixie_cuda_blend_rect(dst.handle, src.handle, px, py, mode))


def blend_rect_masked(dst: DeviceImage, src: DeviceImage, mask: DeviceImage, px, py, mode):
    check(lib().pixie_cuda_blend_rect_masked(dst.handle, src.handle, mask.handle, px, py, mode))


def apply_opacity(image: DeviceImage, opacity):
    check(lib().pixie_cuda_apply_opacity(image.handle, opacity))


def draw(dst: DeviceImage, src: DeviceImage, mat, mode):
    """draw(a, b, transform, blendMode) with any transform (images.nim:636-678)."""
    m = np.ascontiguousarray(mat, np.float32).reshape(9)
    check(lib().pixie_cuda_draw(dst.handle, src.handle, m.ctypes.data, mode))


def draw_tiled(dst: DeviceImage, src: DeviceImage, mat, mode, tiled=True):
    m = np.ascontiguousarray(mat, np.float32).reshape(9)
    fn = lib().pixie_cuda_draw_tiled if tiled else lib().pixie_cuda_draw_correct
    check(fn(dst.handle, src.handle, m.ctypes.data, mode))


def _new_image_from_handle(handle) -> DeviceImage:
    w, h, l, b = i32(0), i32(0), i32(0), i32(0)
    check(lib().pixie_cuda_image_info(handle, C.byref(w), C.byref(h), C.byref(l), C.byref(b), None))
    return DeviceImage(w.value, h.value, l.value, b.value == 1, _handle=handle)


def minify_by2(src: DeviceImage, power=1) -> DeviceImage:
    h = u64(0)
    check(lib().pixie_cuda_minify_by2(src.handle, power, C.byref(h)))
    return _new_image_from_handle(h.value)


def magnify_by2(src: DeviceImage, power=1) -> DeviceImage:
    h = u64(0)
    check(lib().pixie_cuda_magnify_by2(src.handle, power, C.byref(h)))
    return _new_image_from_handle(h.value)


def fill_gradient(image: DeviceImage, kind, handles, stops, opacity=1.0):
    """fillGradient (paints.nim:236-248); stops: [(position, (r, g, b, a))] float colours."""
    hx = np.ascontiguousarray(np.asarray(handles, np.float32).reshape(-1))
    pos = np.ascontiguousarray([s[0] for s in stops], np.float32)
    col = np.ascontiguousarray([s[1] for s in stops], np.float32).reshape(-1)
    check(lib().pixie_cuda_fill_gradient(image.handle, kind, hx.ctypes.data, len(hx) // 2, pos.ctypes.data,
                                         col.ctypes.data, len(stops), opacity))


def fill_gradient_masked(image: DeviceImage, mask: DeviceImage, kind, handles, stops, opacity, mode):
    """Gradient paint composited through a coverage mask in one pass (pixie_cuda_fill_gradient_masked)."""
    hx = np.ascontiguousarray(np.asarray(handles, np.float32).reshape(-1))
    pos = np.ascontiguousarray([s[0] for s in stops], np.float32)
    col = np.ascontiguousarray([s[1] for s in stops], np.float32).reshape(-1)
    check(lib().pixie_cuda_fill_gradient_masked(image.handle, mask.handle, kind, hx.ctypes.data, len(hx) // 2, pos.ctypes.data,
                                                col.ctypes.data, len(stops), opacity, mode))


def blur(image: DeviceImage, lut, radius, oob=0):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_blur(image.handle, lut.ctypes.data, radius, oob))


def blur_rows(image: DeviceImage, lut, radius, oob, y0, y1):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_blur_rows(image.handle, lut.ctypes.data, radius, oob, y0, y1))


def blur_rows_to(src: DeviceImage, dst: DeviceImage, lut, radius, oob, y0, y1):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_blur_rows_to(src.handle, dst.handle, lut.ctypes.data, radius, oob, y0, y1))


def blur_rows_to_flags(src: DeviceImage, dst: DeviceImage, lut, radius, oob, y0, y1, top_flag, bottom_flag, epoch):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_blur_rows_to_flags(src.handle, dst.handle, lut.ctypes.data, radius, oob, y0, y1, top_flag, bottom_flag, epoch))


def blur_rows_x(image: DeviceImage, lut, radius, oob, r0, r1):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_blur_rows_x(image.handle, lut.ctypes.data, radius, oob, r0, r1))


def blur_rows_y(image: DeviceImage, lut, radius, oob, y0, y1):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_blur_rows_y(image.handle, lut.ctypes.data, radius, oob, y0, y1))


def spread_rows(image: DeviceImage, amount, y0, y1):
    check(lib().pixie_cuda_spread_rows(image.handle, amount, y0, y1))


def shadow_rows(src: DeviceImage, dst: DeviceImage, ox, oy, spread_, lut, radius, rgbx, y0, y1):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_shadow_rows(src.handle, dst.handle, ox, oy, spread_, lut.ctypes.data, radius, rgbx, y0, y1))


def spread(image: DeviceImage, amount):
    check(lib().pixie_cuda_spread(image.handle, amount))


def shadow(src: DeviceImage, dst: DeviceImage, ox, oy, spread_, lut, radius, rgbx):
    lut = np.ascontiguousarray(lut, np.uint16)
    check(lib().pixie_cuda_shadow(src.handle, dst.handle, ox, oy, spread_, lut.ctypes.data, radius, rgbx))


class PeerBuffer:
    """A device buffer other processes of the box can map (CUDA IPC), exposed to torch / numpy-style consumers through
    __cuda_array_interface__ (uint8).  `handle` (64 bytes) goes to the neighbours; `PeerBuffer.open(handle, nbytes)`
    maps theirs."""

    def __init__(self, nbytes: int, _ptr=None, _opened=False):
        self.nbytes, self._opened = nbytes, _opened
        if _ptr is not None:
            self.ptr, self.handle = _ptr, None
            return
        p = vp()
        h = (C.c_uint8 * 64)()
        check(lib().pixie_cuda_peer_alloc(nbytes, C.byref(p), h))
        self.ptr, self.handle = p.value, bytes(h)

    @classmethod
    def open(cls, handle: bytes, nbytes: int):
        p = vp()
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        check(lib().pixie_cuda_peer_open(h, C.byref(p)))
        return cls(nbytes, _ptr=p.value, _opened=True)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def close(self):
        if getattr(self, "ptr", None):
            (lib().pixie_cuda_peer_close if self._opened else lib().pixie_cuda_peer_free)(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def halo_push(src_ptr: int, peer_dst_ptr: int, nbytes: int, peer_flag_ptr: int, value: int):
    check(lib().pixie_cuda_halo_push(src_ptr, peer_dst_ptr, nbytes, peer_flag_ptr, value))


class HaloDir(C.Structure):
    """pixie_halo_dir_t of include/pixie_cuda.h."""
    _fields_ = [("src_rows", vp), ("peer_dst_rows", vp), ("bytes", C.c_size_t), ("peer_ready_flag", vp),
                ("peer_data_flag", vp), ("local_ready_flag", vp)]


def halo_exchange(up, down, epoch: int):
    """up / down: HaloDir or None."""
    check(lib().pixie_cuda_halo_exchange(C.byref(up) if up is not None else None, C.byref(down) if down is not None else None, epoch))


def halo_wait2(flag_a, flag_b, value: int):
    check(lib().pixie_cuda_halo_wait2(flag_a, flag_b, value))


def halo_wait(local_flag_ptr: int, value: int):
    check(lib().pixie_cuda_halo_wait(local_flag_ptr, value))


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (for upload_async / download_async)."""

    def __init__(self, nbytes: int):
        p = vp()
        check(lib().pixie_cuda_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, P(C.c_uint8)), (nbytes,))

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.array = None
                lib().pixie_cuda_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def download_async(image: DeviceImage, pinned: PinnedBuffer):
    check(lib().pixie_cuda_image_download_async(image.handle, pinned.ptr))


def upload_async(image: DeviceImage, pinned: PinnedBuffer):
    check(lib().pixie_cuda_image_upload_async(image.handle, pinned.ptr))


PROF_PARTITION, PROF_RASTER, PROF_BLUR_X, PROF_BLUR_Y, PROF_BLEND, PROF_SPREAD, PROF_PLAN = range(7)


def set_profiling(enabled: bool):
    check(lib().pixie_cuda_set_profiling(1 if enabled else 0))


def profile_read(slot: int) -> float:
    v = f32(0)
    check(lib().pixie_cuda_profile_read(slot, C.byref(v)))
    return v.value


def timer_begin():
    check(lib().pixie_cuda_timer_begin())


def timer_end() -> float:
    v = f32(0)
    check(lib().pixie_cuda_timer_end(C.byref(v)))
    return v.value
