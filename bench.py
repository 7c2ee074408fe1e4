#!/usr/bin/env python
"""bench.py — headline benchmark of the raster hot path on B200.

Workload at every N: BASELINE.json configs[1] — the Ghostscript tiger (tests/golden/tiger.svg, the
reference's examples/data/tiger.svg fixture) through parseSvg(data, 4096, 4096) semantics: 227
fills + 78 strokes in document order (90 079 segments), first fill OverwriteBlend, the rest
NormalBlend, into a fresh transparent 4096x4096 RGBX canvas.  One *step* = clear the canvas + one
pass of the ordered command list (partition kernel + fused scanline kernel).  N>1 = one process
per GPU (torchrun), each rank renders its own canvas (independent images, no data-path collective),
scaling "weak".

metric  = Mpixels/s filled+blended: sum over fills of pixels touched with non-zero coverage
          (35.8 Mpx per tiger at 4096^2, counted by the kernel and equal to the oracle's count) / time.
value   = inputs (segments + fill headers) already resident in HBM, CUDA-event timed.
e2e     = through the C-ABI entry point with HOST buffers: pixie_cuda_render_batch_host(host segments in,
          host pixels out, pinned), wall clock, H2D/D2H inside the timed region.
L2 is flushed between timed iterations (256 MiB device write); canvas = 64 MiB < 126 MB L2.

`--impl reference` times the reference's CPU path restated by the oracle (oracle/, kind "port": the
Nim reference cannot be built in this image) on the host cores, same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpixels/s filled+blended (tiger SVG 4096^2)"
UNIT = "Mpixel/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tiger_arrays(size):
    from pixie_b200 import svg as psvg

    with open(os.path.join(ROOT, "tests", "golden", "tiger.svg")) as f:
        data = f.read()
    return psvg.svg_fill_batch(psvg.parseSvg(data, size, size)).arrays()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_tiger(arrays, size, repeats):
    """The oracle (CPU restatement of the reference) rendering the same command list, one thread
    (fills of one canvas are order-dependent and the reference is single-threaded)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import oracle_render_batch

    oracle_render_batch(arrays, size, size)  # warm-up (page faults, library load)
    times, covered = [], 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, covered = oracle_render_batch(arrays, size, size)
        times.append(time.perf_counter() - t0)
    return covered, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g

    g.build_cpu()
    arrays = tiger_arrays(args.size)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import oracle_render_batch

    # N GPUs render N independent canvases: the CPU arm renders N of them too, one host thread each (fills of ONE
    # canvas are order-dependent and the reference is single-threaded, so a canvas cannot use more than one)
    threads = max(1, min(args.gpus, _host_cores()))

    def one_step():
        res = [None] * threads

        def work(k):
            res[k] = oracle_render_batch(arrays, args.size, args.size)[1]

        if threads == 1:
            work(0)
        else:
            _threaded(work, range(threads), threads)
        return sum(res) * args.gpus // threads

    for _ in range(args.warmup):
        one_step()
    times, covered = [], 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        covered = one_step()
        times.append(time.perf_counter() - t0)
    if threads < args.gpus:  # fewer cores than canvases: the remaining canvases would follow in further rounds
        times = [t * args.gpus / threads for t in times]
    ms = 1e3 * sum(times) / len(times)
    val = covered / (ms * 1e-3) / 1e6
    out = {
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "tests/golden/tiger.svg (reference fixture)",
        "config": workload_config(args.size, arrays),
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of {args.gpus} full tiger render(s) at {args.size}^2, one host thread per "
                                   "canvas (fills of a canvas are order-dependent and the reference is single-threaded)"},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def workload_config(size, arrays):
    return {"workload": f"tiger.svg at {size}x{size}: {len(arrays['rgbx'])} ordered fills, "
                        f"{int(arrays['seg_offsets'][-1])} segments, first OverwriteBlend then NormalBlend",
            "canvas": f"{size}x{size} RGBX premultiplied", "l2": "flushed between timed iterations (256 MiB write)",
            "partition": "independent canvas per GPU"}


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _threaded(fn, pieces, cores):
    """Run fn(piece) for every piece on `cores` host threads (the oracle's ctypes calls release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(fn, pieces))


def _bands(rows, parts):
    parts = max(1, min(parts, rows))
    return [(rows * k // parts, rows * (k + 1) // parts) for k in range(parts)]


def _parity(n, bad, mx):
    return {"pixels_compared": int(n), "mismatching": int(bad), "max_abs_delta": int(mx), "fraction": (bad / n if n else None)}


def extras_blend(dev, peak):
    """BASELINE config 3: all 20 modes, 8192^2 dst / src + A8 coverage mask (13 B/px; Overwrite 9 B/px).
    Beside every device time: oracle parity on sampled rows, the oracle's CPU time (1 thread and all host cores, on a
    512-row slab of the same inputs), and for four representative modes the end-to-end time through
    pixie_cuda_blend_rect_masked_host (pinned host pixels in and out)."""
    from pixie_b200 import synth
    from pixie_b200.common import BLEND_MODE_NAMES

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _windows as W
    from _oracle import OracleBackend

    n = 8192
    cores = _host_cores()
    h_dst = np.tile(synth.random_premultiplied(512, n, 0x5EED), (n // 512, 1, 1))
    h_src = np.tile(synth.random_premultiplied(512, n, 0x5EED + 1), (n // 512, 1, 1))
    h_mask = np.tile(synth.coverage_mask(512, n, 0x5EED + 2), (n // 512, 1))
    dst0 = dev.DeviceImage(n, n).upload(h_dst)
    src = dev.DeviceImage(n, n).upload(h_src)
    mask = dev.DeviceImage(n, n, a8=True).upload(h_mask)
    dst = dev.DeviceImage(n, n)
    slab = 512  # CPU sample: rows [0, slab) of the same inputs
    ob = OracleBackend(0)
    pin_d = dev.PinnedBuffer(n * n * 4)
    pin_s = dev.PinnedBuffer(n * n * 4)
    pin_m = dev.PinnedBuffer(n * n)
    pin_s.array[:] = h_src.reshape(-1)
    pin_m.array[:] = h_mask.reshape(-1)
    blends = {}
    for mode in range(20):
        ms = []
        for it in range(4):
            dst.copy_from(dst0)  # also evicts src/mask lines: 3 x 256 MiB planes >> L2
            dev.blend_rect_masked(dst, src, mask, 0, 0, mode)
            t = dev.profile_read(dev.PROF_BLEND)
            if it:
                ms.append(t)
        t = statistics.median(ms)
        nbytes = n * n * (9 if mode == 17 else 13)
        cnt, bad, mx = W.check_blend_rows(dst, h_dst, h_src, h_mask, mode, [(0, 4), (4094, 4098), (n - 4, n)])
        # CPU: the oracle's blendRect on the slab, one thread, then all cores by row bands (blends are order-free)
        work = h_dst[:slab].copy()
        t0 = time.perf_counter()
        ob.blend_rect_masked(work, h_src[:slab], h_mask[:slab], 0, 0, mode)
        t1 = time.perf_counter() - t0
        work[:] = h_dst[:slab]

        def band(b, mode=mode, work=work):
            OracleBackend(0).blend_rect_masked(work[b[0]:b[1]], np.ascontiguousarray(h_src[b[0]:b[1]]),
                                               np.ascontiguousarray(h_mask[b[0]:b[1]]), 0, 0, mode)

        t0 = time.perf_counter()
        _threaded(band, _bands(slab, cores), cores)
        tn = time.perf_counter() - t0
        entry = {"ms": round(t, 4), "GB/s": round(nbytes / t / 1e6, 1), "frac_hbm": round(nbytes / t / 1e6 / peak, 3),
                 "Mpixel/s": round(n * n / t / 1e3, 1), "parity_vs_oracle": _parity(cnt, bad, mx),
                 "cpu_Mpixel/s_1_thread": round(slab * n / t1 / 1e6, 1),
                 f"cpu_Mpixel/s_{cores}_threads": round(slab * n / tn / 1e6, 1)}
        if mode in (0, 7, 3, 12):  # Normal, Overlay, ColorBurn, Hue: end to end through the *_host entry point
            es = []
            for it in range(3):
                pin_d.array[:] = h_dst.reshape(-1)
                t0 = time.perf_counter()
                dev.check(dev.lib().pixie_cuda_blend_rect_masked_host(pin_d.ptr, n, n, pin_s.ptr, pin_m.ptr, 1, n, n, 0, 0, mode))
                es.append(time.perf_counter() - t0)
            te = statistics.median(es[1:])
            got = pin_d.array.reshape(n, n, 4)
            want = h_dst[4094:4098].copy()
            ob.blend_rect_masked(want, np.ascontiguousarray(h_src[4094:4098]), np.ascontiguousarray(h_mask[4094:4098]), 0, 0, mode)
            entry["e2e"] = {"ms": round(te * 1e3, 2), "Mpixel/s": round(n * n / te / 1e6, 1), "h2d_bytes": n * n * 9,
                            "d2h_bytes": n * n * 4, "rows_equal_oracle": bool(np.array_equal(got[4094:4098], want))}
        blends[BLEND_MODE_NAMES[mode]] = entry
    return {"bytes_per_px": "13 (dst r+w, src r, 1-B coverage); Overwrite 9",
            "cpu_baseline": {"kind": "port", "cores": cores,
                             "sample": f"rows [0, {slab}) of the same 8192-wide inputs ({slab * n} px) per mode: the oracle's "
                                       "blendRect + mask composite, 1 thread and all host cores by row bands"},
            "e2e": "pixie_cuda_blend_rect_masked_host, pinned host dst/src/mask in, dst out, wall clock (Normal, Overlay, ColorBurn, Hue)",
            "modes": blends}


def extras_blur_shadow(dev, peak):
    """BASELINE config 4 on one GPU: blur r=32 and the drop shadow on 16384^2 (8 B/px algorithmic), with oracle parity
    through windows (corners, tile seams), the oracle's CPU time on a 2048^2 crop, and end to end through
    pixie_cuda_blur_host / pixie_cuda_shadow_host."""
    from pixie_b200 import host, synth

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _windows as W
    from _oracle import OracleBackend

    out = {}
    n, r = 16384, 32
    cores = _host_cores()
    h_img = np.tile(synth.random_premultiplied(512, n, 0xB10B), (n // 512, 1, 1))
    img = dev.DeviceImage(n, n).upload(h_img)
    work = dev.DeviceImage(n, n)
    lut = host.gaussianKernel(r)
    ms, msk = [], []
    for it in range(5):
        work.copy_from(img)
        dev.timer_begin()
        dev.blur(work, lut, r, 0)
        t = dev.timer_end()
        if it:
            ms.append(t)
            msk.append(dev.profile_read(dev.PROF_BLUR_X))  # the fused kernel (blur_tc.cu) is one launch
    t = statistics.median(ms)
    cnt, bad, mx = W.check_blur_windows(work, h_img, lut, r, 0, W.corner_and_seam_windows(n, n))
    # structured case: an opaque rectangle on a transparent canvas
    h_rect = np.zeros((n, n, 4), np.uint8)
    h_rect[3000:9000, 5000:14000] = (200, 100, 50, 255)
    work.upload(h_rect)
    dev.timer_begin()
    dev.blur(work, lut, r, 0)
    t_rect = dev.timer_end()
    c2, b2, m2 = W.check_blur_windows(work, h_rect, lut, r, 0, [(2976, 3024, 4976, 5024), (8976, 9024, 13976, 14024), (0, 48, 0, 48)])
    del h_rect
    # CPU: the literal oracle (scalar, O(taps) like the reference) on a 2048^2 crop, 1 thread, then all cores by row
    # bands with `radius` halo rows each (every band recomputes the X pass of its halo)
    c = 2048
    crop = np.ascontiguousarray(h_img[:c, :c])
    ob = OracleBackend(0)
    a = crop.copy()
    t0 = time.perf_counter()
    ob.blur(a, lut, r, 0)
    t1 = time.perf_counter() - t0

    def band(b):
        e0, e1 = max(0, b[0] - r), min(c, b[1] + r)
        part = crop[e0:e1].copy()
        OracleBackend(0).blur(part, lut, r, 0)

    t0 = time.perf_counter()
    _threaded(band, _bands(c, cores), cores)
    tn = time.perf_counter() - t0
    # end to end: pinned host pixels in, blurred pixels out
    pin = dev.PinnedBuffer(n * n * 4)
    es = []
    for it in range(3):
        pin.array[:] = h_img.reshape(-1)
        t0 = time.perf_counter()
        dev.check(dev.lib().pixie_cuda_blur_host(pin.ptr, n, n, np.ascontiguousarray(lut, np.uint16).ctypes.data, r, 0))
        es.append(time.perf_counter() - t0)
    te = statistics.median(es[1:])
    got = pin.array.reshape(n, n, 4)
    e2e_ok = bool(np.array_equal(got[8190:8194], work_rows(dev, img, work, lut, r, 8190, 8194)))
    out["blur_r32_16384"] = {
        "ms": round(t, 3), "kernel_ms": round(statistics.median(msk), 3),
        "kernel": "blur_tc_kernel: one fused pass (TMA tiles, tcgen05.mma X and Y contractions, accumulators + ring of X-blurred rows in TMEM)",
        "GB/s": round(n * n * 8 / t / 1e6, 1), "frac_hbm": round(n * n * 8 / t / 1e6 / peak, 3), "bytes_per_px": 8,
        "Mpixel/s": round(n * n / t / 1e3, 1), "Gtaps_per_s": round(n * n * 4 * 2 * 65 / t / 1e6, 1),
        "parity_vs_oracle": _parity(cnt, bad, mx),
        "structured_rect_on_transparent": {"ms": round(t_rect, 3), "parity_vs_oracle": _parity(c2, b2, m2)},
        "cpu_baseline": {"kind": "port", "cores": cores, "Mpixel/s_1_thread": round(c * c / t1 / 1e6, 2),
                         f"Mpixel/s_{cores}_threads": round(c * c / tn / 1e6, 2),
                         "sample": f"the oracle's scalar blur (65 taps x 4 channels x 2 passes per pixel, as the reference) on the "
                                   f"{c}^2 top-left crop of the same image: 1 thread, and {cores} threads by row bands with {r} halo rows"},
        "e2e": {"ms": round(te * 1e3, 2), "Mpixel/s": round(n * n / te / 1e6, 1), "h2d_bytes": n * n * 4, "d2h_bytes": n * n * 4,
                "call": "pixie_cuda_blur_host (pinned host pixels)", "rows_equal_device_path": e2e_ok},
        "traffic_bytes_per_px": "7.9 measured (ncu dram__bytes: 1.09 GB read + 1.03 GB written, profiles/r02_blur_tc_metrics.txt)",
        "bound": "shared-memory port (tensor-pipe operand fetch + plane staging) and worker instruction issue; see DESIGN.md section 4"}
    # ---- shadow: offset (8, 8), spread 4, blur 32, rgba(0, 0, 0, 200)
    h_src = h_img.copy()
    h_src[: n // 3] = 0
    img.upload(h_src)
    col = 0xC8000000
    ms = []
    for it in range(3):
        dev.timer_begin()
        dev.shadow(img, work, 8.0, 8.0, 4, lut, r, col)
        ms.append(dev.timer_end())
    ts = statistics.median(ms[1:])
    cnt, bad, mx = W.check_shadow_windows(work, h_src, (8, 8), 4, lut, r, col,
                                          W.corner_and_seam_windows(n, n, seams=((n // 3, 4096), (8192, 8192 + 32))))
    crop = np.ascontiguousarray(h_src[n // 3 - 1024:n // 3 + 1024, :c])
    t0 = time.perf_counter()
    ob.shadow(crop, 8.0, 8.0, 4, lut, r, col)
    t1 = time.perf_counter() - t0
    pin2 = dev.PinnedBuffer(n * n * 4)
    pin.array[:] = h_src.reshape(-1)
    es = []
    for it in range(2):
        t0 = time.perf_counter()
        dev.check(dev.lib().pixie_cuda_shadow_host(pin.ptr, pin2.ptr, n, n, 8.0, 8.0, 4, np.ascontiguousarray(lut, np.uint16).ctypes.data, r, col))
        es.append(time.perf_counter() - t0)
    te = es[-1]
    got = pin2.array.reshape(n, n, 4)
    e2e_ok = bool(np.array_equal(got[n // 3 - 2:n // 3 + 2], work.download_rows(n // 3 - 2, n // 3 + 2)))
    out["shadow_16384"] = {
        "ms": round(ts, 3), "GB/s": round(n * n * 8 / ts / 1e6, 1), "frac_hbm": round(n * n * 8 / ts / 1e6 / peak, 3),
        "Mpixel/s": round(n * n / ts / 1e3, 1), "parity_vs_oracle": _parity(cnt, bad, mx),
        "cpu_baseline": {"kind": "port", "cores": 1, "Mpixel/s_1_thread": round(crop.shape[0] * crop.shape[1] / t1 / 1e6, 2),
                         "sample": f"the oracle's shadow on a {crop.shape[0]}x{crop.shape[1]} crop across the shape's edge, 1 thread"},
        "e2e": {"ms": round(te * 1e3, 2), "Mpixel/s": round(n * n / te / 1e6, 1), "h2d_bytes": n * n * 4, "d2h_bytes": n * n * 4,
                "call": "pixie_cuda_shadow_host (pinned host pixels)", "rows_equal_device_path": e2e_ok}}
    return out


def work_rows(dev, img, work, lut, r, y0, y1):
    """Rows [y0, y1) of blur(img) on the device path (for the e2e cross-check)."""
    work.copy_from(img)
    dev.blur(work, lut, r, 0)
    return work.download_rows(y0, y1)


def extras_draw_paint(dev, peak):
    """draw with a transform, minifyBy2, gradient fill (8192^2)."""
    from pixie_b200 import host, synth

    n = 8192
    f = np.float32
    dst0 = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 0x5EED), (n // 512, 1, 1)))
    src = dev.DeviceImage(n, n).upload(np.tile(synth.random_premultiplied(512, n, 0x5EED + 1), (n // 512, 1, 1)))
    dst = dev.DeviceImage(n, n)

    def timed(fn):
        ts = []
        for it in range(3):
            dst.copy_from(dst0)
            dev.timer_begin()
            fn()
            t = dev.timer_end()
            if it:
                ts.append(t)
        return statistics.median(ts)

    rot = host.matmul(host.translate(f(n / 2), f(-n / 5)), host.rotate(f(0.5)))
    stops = [(0.0, (1, 0, 0, 1)), (0.3, (0, 1, 0, 0.5)), (1.0, (0, 0, 1, 1))]
    draws = {}
    for name, fn, nbytes in [
        ("draw_rotate_normal", lambda: dev.draw(dst, src, rot, 0), 12 * n * n),
        ("draw_frac_translate_normal", lambda: dev.draw(dst, src, host.translate(f(10.5), f(3.25)), 0), 12 * n * n),
        ("draw_scale_half_normal", lambda: dev.draw(dst, src, host.scale(f(0.5), f(0.5)), 0), 5 * n * n + 3 * n * n),
        ("draw_tiled_scale_0.37", lambda: dev.draw_tiled(dst, src, host.scale(f(0.37), f(0.37)), 0), 12 * n * n),
        ("fill_gradient_linear", lambda: dev.fill_gradient(dst, 3, [(10, 20), (n - 10, n - 30)], stops, 1.0), 4 * n * n),
        ("fill_gradient_radial", lambda: dev.fill_gradient(dst, 4, [(n / 2, n / 2), (n, n / 2), (n / 2, n)], stops, 1.0), 4 * n * n),
        ("fill_gradient_angular", lambda: dev.fill_gradient(dst, 5, [(n / 2, n / 2), (n, n / 2), (n / 2, n)], stops, 1.0), 4 * n * n),
    ]:
        t = timed(fn)
        draws[name] = {"ms": round(t, 4), "GB/s": round(nbytes / t / 1e6, 1), "frac_hbm": round(nbytes / t / 1e6 / peak, 3)}
    # non-solid paint composite (paths.nim:2115-2142) through the coverage mask of an ellipse covering ~21 % of the canvas:
    # the gradient evaluated inside the masked blend against fillGradient + blend_rect_masked
    mask = dev.DeviceImage(n, n)
    ell = host.newPath()
    ell.ellipse(n / 2, n / 2, n * 0.3, n * 0.22)
    dev.fill_segments(mask, host.fill_segments(ell), 0xFFFFFFFF, 0, 0)
    fill = dev.DeviceImage(n, n)
    hl, hr = [(n * 0.2, n * 0.3), (n * 0.9, n * 0.7)], [(n / 2, n / 2), (n * 0.9, n / 2), (n / 2, n * 0.95)]
    paints = {}
    for name, kind, handles in (("linear", 3, hl), ("radial", 4, hr), ("angular", 5, hr)):
        def separate():
            dev.fill_gradient(fill, kind, handles, stops, 1.0)
            dev.blend_rect_masked(dst, fill, mask, 0, 0, 0)

        t_sep = timed(separate)
        want = dst.download_rows(n // 2 - 8, n // 2 + 8)
        t_fused = timed(lambda: dev.fill_gradient_masked(dst, mask, kind, handles, stops, 1.0, 0))
        same = bool(np.array_equal(want, dst.download_rows(n // 2 - 8, n // 2 + 8)))
        paints[name] = {"fused_ms": round(t_fused, 4), "separate_passes_ms": round(t_sep, 4), "rows_equal": same}
    return {"bytes_per_px": "draw 12 (dst r+w, src r); scale 0.5: minify 5/src px + draw over the covered quarter; "
                            "gradient 4 (write)", "ops": draws,
            "gradient_paint_composite_normal": {"what": "fillPath with a gradient paint, mask given: pixie_cuda_fill_gradient_masked "
                                                        "(gradient evaluated inside the masked blend, mask-0 pixels skipped) vs "
                                                        "fillGradient + blend_rect_masked; ellipse covering 21 % of 8192^2",
                                                "paints": paints}}


def extras_flatten(dev):
    """the tiger from path COMMANDS — flattening, stroking and shapesToSegments on the device
    (pixie_cuda_cmdlist_create_from_paths) against libpixie_host.so (the C++ mirror of the reference's host code) +
    pixie_cuda_cmdlist_create from its segments.  Wall clock, inputs in host memory, the list resident afterwards."""
    from pixie_b200 import host, svg as psvg

    size = 4096
    doc = psvg.parseSvg(open(os.path.join(ROOT, "tests", "golden", "tiger.svg")).read(), size, size)
    calls = []
    for d, props in doc.elements:
        if not (props.display and props.opacity > 0):
            continue
        path = host.parsePath(d)
        if props.fill != "none":
            calls.append((host.fill_segments, (path, props.transform)))
        if props.stroke != 0 and props.strokeWidth > 0:
            calls.append((host.stroke_segments, (path, props.transform, props.strokeWidth, props.strokeLineCap, props.strokeLineJoin,
                                                props.strokeMiterLimit, props.strokeDashArray)))
    tc = []
    for it in range(4):
        t0 = time.perf_counter()
        for fn, a in calls:
            fn(*a)
        tc.append(time.perf_counter() - t0)
    arrays = psvg.svg_fill_batch(doc).arrays()
    pb = psvg.svg_path_batch(doc)
    packed = pb.packed()
    th, td = [], []
    for it in range(6):
        dev.sync()
        t0 = time.perf_counter()
        cl = dev.CmdList(size, size, 1, arrays)
        dev.sync()
        th.append(time.perf_counter() - t0)
        del cl
        t0 = time.perf_counter()
        cl = dev.CmdList.from_paths(size, size, 1, pb, packed)
        dev.sync()
        td.append(time.perf_counter() - t0)
        if it == 5:
            xy, wd, so = cl.segments()
            same = bool(np.array_equal(xy, arrays["xyxy"]) and np.array_equal(wd, arrays["winding"]) and
                        np.array_equal(so, arrays["seg_offsets"]))
        del cl
    # end to end, host pixels out: commands -> pixels in one call against host flattening + pixie_cuda_render_batch_host
    pinned = dev.PinnedBuffer(size * size * 4)
    te_dev, te_host = [], []
    for it in range(5):
        dev.sync()
        t0 = time.perf_counter()
        dev.render_paths_host(pinned.ptr, size, size, pb, packed)
        te_dev.append(time.perf_counter() - t0)
        if it == 4:
            px_paths = pinned.array[:size * size * 4].copy()
        t0 = time.perf_counter()
        dev.render_batch_host(pinned.ptr, size, size, arrays)
        te_host.append(time.perf_counter() - t0)
    same_px = bool(np.array_equal(px_paths, pinned.array[:size * size * 4]))
    t_host = statistics.median(tc[1:]) + statistics.median(th[2:])
    t_dev = statistics.median(td[2:])
    ncmd = sum(d.num_commands for d in pb.descs)
    return {"paths": len(pb), "commands": ncmd, "segments": int(len(arrays["winding"])), "paths_flattened_on_host": pb.host_paths,
            "h2d_bytes": {"commands_and_headers": int(len(packed[1]) * 4 + len(pb) * 80), "segments_host_path": int(len(arrays["winding"]) * 18)},
            "device": {"ms": round(t_dev * 1e3, 3), "call": "pixie_cuda_cmdlist_create_from_paths (H2D of the command stream, resolve / count / "
                                                            "scan / emit / stroke / bounds kernels, 3 small readbacks, list build)"},
            "cpu_baseline": {"kind": "port", "cores": 1, "ms": round(t_host * 1e3, 3),
                             "flatten_ms": round(statistics.median(tc[1:]) * 1e3, 3), "cmdlist_create_ms": round(statistics.median(th[2:]) * 1e3, 3),
                             "sample": f"{len(calls)} fill_segments / stroke_segments calls into libpixie_host.so (commandsToShapes + "
                                       "strokeShapes + shapesToSegments as the reference's host code, 1 thread) + pixie_cuda_cmdlist_create"},
            "speedup": round(t_host / t_dev, 1), "segments_equal_host_flattener": same,
            "e2e_commands_to_host_pixels": {
                "ms": round(statistics.median(te_dev[1:]) * 1e3, 3), "call": "pixie_cuda_render_paths_host (87 KB of commands in, 64 MiB of pinned pixels out)",
                "host_flatten_path_ms": round((statistics.median(tc[1:]) + statistics.median(te_host[1:])) * 1e3, 3),
                "host_flatten_path": "libpixie_host.so flattening + pixie_cuda_render_batch_host", "pixels_equal": same_px}}


def icons_batch(dev, rank, world, n_icons=1024, size=512, cpu_sample=128):
    """BASELINE config 5 (scaled to fit the time budget): synthetic icons, 512^2 each, every icon its own
    layer of one device allocation, this rank's contiguous shard, ONE launch set for the whole shard;
    results stay on the device (checksum), algorithmic bytes 8 B x covered px + 18 B x segments.
    `ms` = device time of one run of the resident list (count pass included); `e2e` = wall clock from host segment
    arrays to the on-device checksum: command-list creation (H2D of the segments, count / scan kernels, readback of
    the sizes) + the run + the checksum's 8-byte D2H.  CPU: the oracle over a sample of the same icons."""
    from pixie_b200 import multi, synth
    from pixie_b200.device import FillBatch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import oracle_render_batch

    b0, b1 = multi.shard_range(n_icons * world, world, rank)
    batch = FillBatch()
    per_icon = []
    for i in range(b0, b1):
        one = FillBatch()
        synth.icon_fills(i, size, 0, one)
        if i - b0 < cpu_sample:
            per_icon.append(one.arrays())
        synth.icon_fills(i, size, i - b0, batch)
    arrays = batch.arrays()
    img = dev.DeviceImage(size, size, b1 - b0)
    cl = dev.CmdList(size, size, b1 - b0, arrays)
    covered = cl.run(img, count_covered=True)
    ms = []
    for it in range(4):
        img.fill(0)
        dev.timer_begin()
        cl.run(img)
        t = dev.timer_end()
        if it:
            ms.append(t)
    t = statistics.median(ms)
    nseg = int(arrays["seg_offsets"][-1])
    checksum = img.checksum()
    es = []
    for it in range(3):
        img.fill(0)
        dev.sync()
        t0 = time.perf_counter()
        cl2 = dev.CmdList(size, size, b1 - b0, arrays)
        cl2.run(img)
        c2 = img.checksum()
        es.append(time.perf_counter() - t0)
        del cl2
    te = statistics.median(es[1:])
    # the same icons from path COMMANDS: flattening + stroking on the device (round joins / caps go through the host)
    from pixie_b200.device import PathBatch

    pb = PathBatch()
    for i in range(b0, b1):
        synth.icon_fills(i, size, i - b0, paths=pb)
    packed = pb.packed()
    ep = []
    for it in range(3):
        img.fill(0)
        dev.sync()
        t0 = time.perf_counter()
        cl3 = dev.CmdList.from_paths(size, size, b1 - b0, pb, packed)
        cl3.run(img)
        c3 = img.checksum()
        ep.append(time.perf_counter() - t0)
        del cl3
    tp = statistics.median(ep[1:])
    from_paths = {"ms": round(tp * 1e3, 3), "icons_per_s": round((b1 - b0) / tp), "paths": len(pb), "paths_flattened_on_host": pb.host_paths,
                  "h2d_bytes": int(len(packed[1]) * 4 + len(pb) * 80 + len(packed[3]) * 18),
                  "call": "pixie_cuda_cmdlist_create_from_paths + _run + _image_checksum from the icons' path commands",
                  "checksum_equal": bool(c3 == checksum)}
    # parity + CPU baseline: the oracle renders the first `cpu_sample` icons (independent canvases: one icon per
    # thread task), the GPU's layers of the same icons must be byte-equal
    cores = _host_cores()
    results = [None] * len(per_icon)

    def one_icon(k):
        results[k] = oracle_render_batch(per_icon[k], size, size)

    one_icon(0)
    t0 = time.perf_counter()
    _threaded(one_icon, range(len(per_icon)), cores)
    tc = time.perf_counter() - t0
    cov_cpu = sum(r[1] for r in results)
    layers = img.download().reshape(b1 - b0, size, size, 4)
    bad = sum(int((layers[k] != results[k][0][0]).any(axis=-1).sum()) for k in range(len(per_icon)))
    return {"icons": b1 - b0, "fills": len(arrays["rgbx"]), "segments": nseg, "covered_px": int(covered), "ms": round(t, 3),
            "Mpixel/s": round(covered / t / 1e3, 1), "icons_per_s": round((b1 - b0) / t * 1e3),
            "GB/s_algorithmic": round((8 * covered + 18 * nseg) / t / 1e6, 1), "checksum": checksum,
            "e2e": {"ms": round(te * 1e3, 3), "icons_per_s": round((b1 - b0) / te), "Mpixel/s": round(covered / te / 1e6, 1),
                    "h2d_bytes": nseg * 18 + len(arrays["rgbx"]) * 56, "d2h_bytes": 8 + 24,
                    "call": "pixie_cuda_cmdlist_create + _run + _image_checksum from host segment arrays",
                    "checksum_equal": bool(c2 == checksum)},
            "e2e_from_path_commands": from_paths,
            "parity_vs_oracle": _parity(len(per_icon) * size * size, bad, 0 if bad == 0 else 255),
            "cpu_baseline": {"kind": "port", "cores": cores, "icons_per_s": round(len(per_icon) / tc, 1),
                             "Mpixel/s": round(cov_cpu / tc / 1e6, 2),
                             "sample": f"the first {len(per_icon)} icons of the same shard rendered by the oracle, one icon per task on {cores} host threads"}}


def icons_sharded(dev, dist, rank, world, per_rank, size=512, chunk=2048):
    """BASELINE config 5 at scale: `per_rank * world` synthetic 512^2 icons (100 000 on 8 GPUs), contiguous shard per
    rank (multi.shard_range), rendered chunk by chunk into one reused stack of canvases; results stay on the device
    (checksum of checksums), no data-path collective.  Time = device time of the launches, max over ranks."""
    import torch

    from pixie_b200 import multi, synth
    from pixie_b200.device import FillBatch

    total = per_rank * world
    b0, b1 = multi.shard_range(total, world, rank)
    img = dev.DeviceImage(size, size, min(chunk, b1 - b0))
    ms, wall, covered, nseg, nfills, checksum = 0.0, 0.0, 0, 0, 0, 0
    for c0 in range(b0, b1, chunk):
        c1 = min(c0 + chunk, b1)
        batch = FillBatch()
        for i in range(c0, c1):
            synth.icon_fills(i, size, i - c0, batch)  # host path synthesis + flattening: outside both timers
        arrays = batch.arrays()
        if c1 - c0 != img.layers:
            img = dev.DeviceImage(size, size, c1 - c0)
        img.fill(0)
        dev.sync()
        # wall clock from host segment arrays to the on-device checksum: command-list creation (bounds pass, H2D of the
        # segments, count kernels + readback of the sizes), the run, the checksum's 8-byte D2H
        t0 = time.perf_counter()
        cl = dev.CmdList(size, size, c1 - c0, arrays)
        dev.timer_begin()
        covered += cl.run(img, count_covered=True)
        ms += dev.timer_end()
        cs = img.checksum()
        wall += time.perf_counter() - t0
        nseg += int(arrays["seg_offsets"][-1])
        nfills += len(arrays["rgbx"])
        checksum = (checksum * 1000003 + cs) % (1 << 64)
        del cl
    t = torch.tensor([ms, float(covered), float(nseg), float(nfills), wall * 1e3], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_max, wall_max = float(tmax[0].item()), float(tmax[4].item())
    cov, segs = float(t[1].item()), float(t[2].item())
    return {"icons": total, "icons_per_rank": b1 - b0, "fills": int(t[3].item()), "segments": int(segs), "covered_px": int(cov),
            "ms_max_over_ranks": round(ms_max, 3), "icons_per_s": round(total / ms_max * 1e3),
            "Mpixel/s": round(cov / ms_max / 1e3, 1), "GB/s_algorithmic": round((8 * cov + 18 * segs) / ms_max / 1e6, 1),
            "e2e": {"ms_max_over_ranks": round(wall_max, 3), "icons_per_s": round(total / wall_max * 1e3),
                    "call": "pixie_cuda_cmdlist_create (bounds, H2D of the segments, count pass) + _run + _image_checksum "
                            "per chunk from host segment arrays, wall clock; host path synthesis excluded"},
            "checksum_rank0": checksum, "scaling": "weak (12 500 icons per GPU)", "chunk": chunk}


def _band_tile(n, k, rows=256):
    """Rank k's band of the synthetic 16384-wide canvas: a 256-row tile of noise (seeded by the rank) repeated."""
    from pixie_b200 import synth

    return synth.random_premultiplied(rows, n, 0xB10B + k)


def _global_rows(n, world, y0, y1, transparent_above=0):
    """Rows [y0, y1) of the global canvas assembled on the host from the per-rank tiles (for the oracle windows)."""
    from pixie_b200 import multi

    out = np.zeros((y1 - y0, n, 4), np.uint8)
    for k in range(world):
        a, b = multi.band_range(n, world, k)
        lo, hi = max(a, y0), min(b, y1)
        if lo >= hi:
            continue
        tile = _band_tile(n, k)
        idx = (np.arange(lo, hi) - a) % tile.shape[0]
        out[lo - y0:hi - y0] = tile[idx]
    if transparent_above > y0:
        out[: min(transparent_above, y1) - y0] = 0
    return out


class _BandView:
    """download_rows over a band tensor in global row coordinates (what tests/_windows.py reads)."""

    def __init__(self, t, y0):
        self.t, self.y0 = t, y0

    def download_rows(self, a, b):
        return self.t[a - self.y0:b - self.y0].cpu().numpy()


def host_link_probe(dev, dist, world, mib=64, reps=10):
    """What the e2e number runs into at N > 1: every rank returns 64 MiB of pixels per step over its PCIe link.  All
    ranks copy a 64 MiB device buffer to pinned host memory at the same time; the aggregate is the host side's ceiling
    for N concurrent result streams (root complexes / host memory), measured in the same run."""
    import torch

    n = mib << 20
    img = dev.DeviceImage(n // 4 // 4096, 4096)
    pin = dev.PinnedBuffer(n)
    for _ in range(2):
        dev.download_async(img, pin)
    dev.sync()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.download_async(img, pin)
    dev.sync()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tmax = float(t.item())
    return {"d2h_GBps_per_rank_concurrent": round(n * reps / tmax / 1e9, 1), "d2h_GBps_aggregate": round(world * n * reps / tmax / 1e9, 1),
            "MiB_per_copy": mib, "ranks_copying_at_once": world}


def row_bands_multi_gpu(dev, dist, rank, world, local_rank, peak):
    """BASELINE config 4 (+ the other row-band rows of SURVEY 8e) across N GPUs: one canvas in row bands.
    blur / spread / shadow exchange halo rows with ncclSend / ncclRecv (torch.distributed P2P over NVLink) — for the
    blur the exchange runs on a side stream while the halo-free horizontal pass blurs the band's own rows; blends and
    fills need no exchange.  Every result is compared with the ORACLE on windows of this rank's band (two of them
    straddling the cuts), built from the global canvas on the host."""
    import torch

    from pixie_b200 import host, multi, svg as psvg
    from pixie_b200.common import NormalBlend

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _windows as W
    from _oracle import OracleBackend

    out = {}
    n, r = 16384, 32
    y0, y1 = multi.band_range(n, world, rank)
    lut = host.gaussianKernel(r)

    def fill_band(rb, transparent_above=0):
        tile = torch.from_numpy(_band_tile(n, rank)).cuda()
        rb.band.copy_(tile.repeat((y1 - y0 + 255) // 256, 1, 1)[: y1 - y0])
        if transparent_above > y0:
            rb.band[: min(transparent_above, y1) - y0] = 0

    def timed(fn, reps=5):
        ts = []
        for it in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if it:
                ts.append(e0.elapsed_time(e1))
        t = torch.tensor([statistics.median(ts)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(bad):
        t = torch.tensor([int(bad)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def windows():  # this rank's band: across its upper cut, across its lower cut, interior; all 48 x 48
        wins = [(y0, y0 + 48, 0, 48), (y1 - 48, y1, n - 48, n), ((y0 + y1) // 2, (y0 + y1) // 2 + 48, 8192 - 24, 8192 + 24)]
        return wins

    def crop_check(fn_check, band_tensor, reach, transparent_above=0):
        bad = cnt = 0
        view = _BandView(band_tensor, y0)
        for win in windows():
            a, b = max(0, win[0] - reach - 8), min(n, win[1] + reach + 8)
            rows = _global_rows(n, world, a, b, transparent_above)
            # windows in the coordinates of the row slab [a, b): the slab's top / bottom are image borders only where
            # a == 0 / b == n, otherwise they are farther than `reach` from the window
            sub = (win[0] - a, win[1] - a, win[2], win[3])
            c, bd, _ = fn_check(_BandView(band_tensor, y0 - a), rows, [sub])
            cnt += c
            bad += bd
        return cnt, bad

    margin = 64
    rb = multi.RowBand(n, n, rank, world, margin=margin)
    # ---- blur r=32: exchange overlapped with the horizontal pass, and the plain exchange-then-blur for comparison
    # (timed on the band as it is: the cost does not depend on the pixel values; refilled before the parity run)
    fill_band(rb)
    t_ov = timed(lambda: rb.blur(r, lut, 0, overlap=True))
    t_se = timed(lambda: rb.blur(r, lut, 0, overlap=False))
    fill_band(rb)
    rb.blur(r, lut, 0, overlap=True)
    cnt, bad = crop_check(lambda v, rows, wins: W.check_blur_windows(v, rows, lut, r, 0, wins), rb.band, r)
    out["blur_r32_16384_row_bands"] = {
        "ms": round(t_ov, 3), "ms_exchange_then_blur": round(t_se, 3), "GB/s": round(n * n * 8 / t_ov / 1e6, 1),
        "frac_hbm_aggregate": round(n * n * 8 / t_ov / 1e6 / (peak * world), 3),
        "parity_vs_oracle": _parity(all_ok(cnt), all_ok(bad), 0), "halo_bytes_per_interior_edge": 2 * r * n * 4,
        "scaling": "strong (one 16384^2 canvas)", "transport": rb.transport,
        "exchange": "peer: a kernel stores the band's edge rows into the neighbour's CUDA-IPC mapped margin over NVLink and "
                    "publishes an epoch flag there (pixie_cuda_halo_push / _wait, no NCCL launch); the band's interior rows "
                    "(no halo needed) are blurred meanwhile, the two edge strips after the flag; ms_exchange_then_blur = "
                    "the same without the split"}
    # ---- spread 4 and the drop shadow (offset (8, 8), spread 4, blur 32): halo = |offset.y| + |spread| + radius rows
    fill_band(rb)
    t_sp = timed(lambda: rb.spread(4))
    fill_band(rb)
    rb.spread(4)
    cnt, bad = crop_check(lambda v, rows, wins: W.check_spread_windows(v, rows, 4, wins), rb.band, 4)
    out["spread4_16384_row_bands"] = {"ms": round(t_sp, 3), "parity_vs_oracle": _parity(all_ok(cnt), all_ok(bad), 0)}
    ta = n // 3
    fill_band(rb, transparent_above=ta)
    col = 0xC8000000
    t_sh = timed(lambda: rb.shadow((8.0, 8.0), 4, r, lut, col), reps=4)
    sh = rb.shadow((8.0, 8.0), 4, r, lut, col)
    cnt, bad = crop_check(lambda v, rows, wins: W.check_shadow_windows(v, rows, (8, 8), 4, lut, r, col, wins), sh, 8 + 4 + r, ta)
    out["shadow_16384_row_bands"] = {"ms": round(t_sh, 3), "GB/s": round(n * n * 8 / t_sh / 1e6, 1),
                                     "halo_rows": 8 + 4 + r, "parity_vs_oracle": _parity(all_ok(cnt), all_ok(bad), 0)}
    del rb, sh
    # ---- blends in row bands (no exchange): NormalBlend of a 16384-wide src band over the dst band
    rbd = multi.RowBand(n, n, rank, world, margin=0)
    fill_band(rbd)
    src = torch.from_numpy(_band_tile(n, rank + 100)).cuda().repeat((y1 - y0 + 255) // 256, 1, 1)[: y1 - y0].contiguous()
    t_bl = timed(lambda: rbd.blend(src, NormalBlend))
    fill_band(rbd)
    rbd.blend(src, NormalBlend)
    want = _global_rows(n, world, y0, y0 + 4)
    OracleBackend(0).blend_rect(want, np.ascontiguousarray(_band_tile(n, rank + 100)[:4]), 0, 0, NormalBlend)
    bad = int((rbd.band[:4].cpu().numpy() != want).any(axis=-1).sum())
    out["blend_normal_16384_row_bands"] = {"ms": round(t_bl, 3), "GB/s": round(n * n * 12 / t_bl / 1e6, 1),
                                           "frac_hbm_aggregate": round(n * n * 12 / t_bl / 1e6 / (peak * world), 3),
                                           "parity_vs_oracle": _parity(all_ok(4 * n), all_ok(bad), 0), "exchange": "none (per-pixel)"}
    del rbd, src
    # ---- one canvas of fills in row bands: the tiger at 8192^2, every rank holds the whole command list (replicated
    # upload) and plans + rasterises only its rows; partition boundaries are those of the whole canvas
    size = 8192
    with open(os.path.join(ROOT, "tests", "golden", "tiger.svg")) as f:
        arrays = psvg.svg_fill_batch(psvg.parseSvg(f.read(), size, size)).arrays()
    fy0, fy1 = multi.band_range(size, world, rank)
    rbf = multi.RowBand(size, size, rank, world, margin=0)
    cl = dev.CmdList(size, size, 1, arrays)

    def fills():
        rbf.band.zero_()
        return rbf.fill(cl)

    t_fi = timed(fills)
    rbf.band.zero_()
    covered = rbf.fill(cl, count_covered=True)
    from _util import oracle_render_batch

    want_full, cov_cpu = oracle_render_batch(arrays, size, size)
    bad = int((rbf.band.cpu().numpy() != want_full[0][fy0:fy1]).any(axis=-1).sum())
    tcov = torch.tensor([float(covered)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tcov, op=dist.ReduceOp.SUM)
    out["tiger_8192_fills_row_bands"] = {"ms": round(t_fi, 3), "Mpixel/s": round(float(tcov.item()) / t_fi / 1e3, 1),
                                         "covered_px_sum_over_ranks": int(tcov.item()), "covered_px_oracle": int(cov_cpu),
                                         "parity_vs_oracle": _parity(all_ok((fy1 - fy0) * size), all_ok(bad), 0),
                                         "exchange": "none (segment list replicated; global partition boundaries, band-restricted plan + raster)"}
    dev.set_stream(None)
    return out


def run_ours(args):
    import __graft_entry__ as g
    from pixie_b200 import device as dev

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_

        torch.cuda.set_device(local_rank)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_
    if not os.path.exists(dev.LIB_PATH):
        g.build()
    numa = None
    if world > 1:
        from pixie_b200 import multi

        numa = multi.bind_to_gpu_numa(local_rank)  # before any pinned allocation
    dev.init(local_rank)
    dev.set_profiling(True)
    peak, peak_src = load_peaks()
    size = args.size
    arrays = tiger_arrays(size)
    nseg = int(arrays["seg_offsets"][-1])

    def barrier():
        dev.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    img = dev.DeviceImage(size, size)
    flush = dev.DeviceImage(8192, 8192)  # 256 MiB > 126 MB L2
    cl = dev.CmdList(size, size, 1, arrays)
    covered = cl.run(img, count_covered=True)
    info = cl.info()
    pinned = dev.PinnedBuffer(size * size * 4)

    def l2_flush(i):
        flush.fill(0x01020304 + i)

    # ---- device-resident timing (value)
    for i in range(args.warmup):
        l2_flush(i)
        cl.run(img, clear=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, part_ms, plan_ms, rast_ms = [], [], [], []
    launches0 = dev.launch_count()
    launches_timed = 0
    barrier()
    for i in range(args.steps):
        l2_flush(i)
        dev.sync()
        l0 = dev.launch_count()
        dev.timer_begin()
        cl.run(img, clear=True)  # newImage + fills: the raster kernel zeroes every row tile before its first fill
        step_ms.append(dev.timer_end())
        launches_timed += dev.launch_count() - l0
        part_ms.append(dev.profile_read(dev.PROF_PARTITION))
        plan_ms.append(dev.profile_read(dev.PROF_PLAN))
        rast_ms.append(dev.profile_read(dev.PROF_RASTER))
    barrier()
    total_ms = max_over_ranks(sum(step_ms))
    ms_per_step = total_ms / args.steps
    value = world * covered / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers
    for i in range(min(args.warmup, 3)):
        dev.render_batch_host(pinned.ptr, size, size, arrays)
    e2e_s = []
    barrier()
    for i in range(args.steps):
        l2_flush(i)
        dev.sync()
        t0 = time.perf_counter()
        # one C-ABI call: fresh canvas, host segments -> H2D -> count/scan/partition/raster kernels -> canvas -> pinned host
        dev.render_batch_host(pinned.ptr, size, size, arrays)
        e2e_s.append(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()
    e2e_total = max_over_ranks(sum(e2e_s))
    e2e_value = world * covered / (e2e_total / args.steps) / 1e6
    h2d = nseg * 18 + len(arrays["rgbx"]) * 56 + (info["partitions"] * 2 + 1) * 4
    d2h = size * size * 4
    checksum_ok = int(pinned.array[:size * size * 4].view(np.uint32).sum(dtype=np.uint64)) != 0

    banded = None
    icons_multi = None
    if dist is not None and not args.no_extras:
        try:
            banded = row_bands_multi_gpu(dev, dist, rank, world, local_rank, peak)
        except Exception as e:
            import traceback

            banded = {"error": repr(e), "trace": traceback.format_exc()[-800:]}
        try:
            icons_multi = icons_sharded(dev, dist, rank, world, args.icons_per_gpu)
        except Exception as e:
            icons_multi = {"error": repr(e)}
    host_link = None
    if dist is not None and not args.no_extras:
        try:
            host_link = host_link_probe(dev, dist, world)
        except Exception as e:
            host_link = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (raster_kernel), algorithmic bytes per SURVEY.md 8(d)
    alg_bytes = 8 * covered + 18 * nseg
    rast = statistics.mean(rast_ms)
    part = statistics.mean(part_ms)
    achieved = alg_bytes / (rast * 1e-3) / 1e9
    roofline = {"kernel": "raster_kernel", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of one raster_kernel launch on this workload, from the offline
                # `ncu --set full` capture summarised in profiles/r02_tiger_metrics.txt (18.07 MB read + 34.84 MB written, the
                # latter incl. what reaches DRAM of the 64 MiB the launch clears: the canvas stays in the 126 MB L2) — a
                # capture, not measured by this run (traffic_source says so)
                "traffic": 52912896 if size == 4096 else None, "traffic_source": "profiles/r02_tiger_metrics.txt (ncu capture)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": round(rast, 4),
                "partition_kernel_ms": round(part, 4), "plan_kernel_ms": round(statistics.mean(plan_ms), 4), "share_of_step": round(rast / statistics.mean(step_ms), 3),
                # since round 2 the launch also clears the canvas (pixie_cuda_cmdlist_run_cleared: every row tile is zeroed by the
                # warp that rasterises it): 4 B per canvas pixel that `achieved` (SURVEY.md 8(d)'s per-unit figure) does not count
                "clear_bytes_per_launch": 4 * size * size,
                "achieved_incl_clear": round((alg_bytes + 4 * size * size) / (rast * 1e-3) / 1e9, 1),
                "note": "305 order-dependent fills: the kernel ends with its longest (row, tile) ticket — latency-bound, not "
                        "bandwidth-bound (SURVEY.md 7.1; per-ticket cycle counters: DESIGN.md 8)"}

    cpu = None
    extras = None
    if world == 1:
        g.build_cpu()
        reps = max(3, min(40, int(args.cpu_seconds / 0.35)))
        ccov, times = cpu_tiger(arrays, size, reps)
        cval = ccov / statistics.mean(times) / 1e6
        cpu = {"value": round(cval, 3), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{reps} full tiger renders at {size}^2 with the oracle (same workload, 1 thread: fills of a "
                         "canvas are order-dependent; the reference is single-threaded)",
               "covered_px_equal_to_gpu": bool(ccov == covered)}
        if not args.no_extras:
            extras = {}
            for key, fn in [("blend_8192_masked", lambda: extras_blend(dev, peak)), ("blur_shadow_16384", lambda: extras_blur_shadow(dev, peak)),
                            ("draw_paint_8192", lambda: extras_draw_paint(dev, peak)), ("icons_512_batch", lambda: icons_batch(dev, 0, 1)),
                            ("tiger_from_path_commands_4096", lambda: extras_flatten(dev))]:
                try:  # extras never block the headline line
                    extras[key] = fn()
                except Exception as e:
                    import traceback

                    extras[key] = {"error": repr(e), "trace": traceback.format_exc()[-600:]}

    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "tests/golden/tiger.svg (reference fixture); synthetic for extras",
        "config": workload_config(size, arrays), "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(1e3 * e2e_total / args.steps, 4), "result_nonzero": checksum_ok},
        "gpu_launches": int(launches_timed),
        "covered_px_per_step": int(covered), "cmdlist": info, "numa_binding_rank0": numa,
        "roofline": roofline, "cpu_baseline": cpu,
    }
    if banded is not None:
        extras = dict(extras or {})
        extras["row_bands"] = banded
        extras["icons_512_sharded"] = icons_multi
        extras["host_link"] = host_link
    if extras is not None:
        out["extras"] = extras
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--icons-per-gpu", type=int, default=12500, help="N>1 extras: icons per GPU (BASELINE config 5: 100k on 8)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
