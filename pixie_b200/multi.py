"""One process per GPU: how the hot path shards across the GPUs of one B200 box.

* independent images (icons, per-rank canvases): ``shard_range`` — contiguous blocks of units per
  rank, no data-path collective;
* one giant canvas: horizontal row bands (``band_range``).  Fills and blends are per-pixel/per-row
  independent, so bands need no exchange.  ``blur`` / ``spread`` / ``shadow`` need `radius` rows from
  the vertical neighbours: ``exchange_halos`` moves them with point-to-point send/recv
  (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests) and ``blur_band`` then runs
  the row-band form of the kernel (pixie_cuda_blur_rows) on band + halo.  The halo carries *input*
  rows; the X pass is recomputed on them, which gives the same bytes as exchanging the X-blurred
  intermediate because that intermediate is quantised per row (images.nim:341-352).
"""
from __future__ import annotations

import numpy as np


def bind_to_gpu_numa(device_index: int) -> dict:
    """One process per GPU: run this process on the CPUs of the NUMA node its GPU hangs off, so that the pinned
    staging buffers (first touch) and the host passes of the end-to-end path stay local to the PCIe root the
    pixels cross.  Best effort: returns what it found and changes nothing when the topology is not exposed."""
    import os

    info = {"numa_node": None, "cpus": None}
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml reports an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # no NVML, no sysfs entry, container without the topology: leave the affinity alone
        info["error"] = repr(e)[:80]
    return info


def torch_stream_handle() -> int:
    """The CUDA stream torch is issuing on, as a handle pixie_cuda_set_stream understands.  torch's default stream is
    the legacy stream, whose handle is 0 — which the C ABI reads as "the library's own stream"; cudaStreamLegacy (1)
    names it explicitly, so that the kernels are ordered with torch's and NCCL's work on that stream."""
    import torch

    return torch.cuda.current_stream().cuda_stream or 1


def shard_range(n_units: int, world: int, rank: int):
    """Contiguous block [begin, end) of independent units owned by `rank`."""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def band_range(height: int, world: int, rank: int):
    """Row band [y0, y1) of a canvas of `height` rows owned by `rank`."""
    return shard_range(height, world, rank)


def check_halo_reach(sizes, radius: int):
    """A halo of `radius` rows must come from the direct neighbour alone.  `sizes` (all band heights) is the same on
    every rank, so every rank takes the same decision: either all raise or none does — a rank that raised while
    its neighbours entered the exchange would leave them waiting for a peer that never posts."""
    world = len(sizes)
    for k in range(1, world - 1):  # interior bands feed both neighbours
        if sizes[k] < radius:
            raise ValueError("bands shorter than the filter's reach need multi-hop halos: use fewer ranks")


def exchange_halos(band, radius: int, rank: int, world: int, group=None):
    """band: torch uint8 tensor [rows, width, 4] (this rank's rows).  Returns (ext, top, bottom):
    ext = [halo_top ; band ; halo_bottom] with `top` / `bottom` halo rows actually received from the
    neighbours (0 at the image border, and at most the neighbour's band height)."""
    import torch
    import torch.distributed as dist

    rows = band.shape[0]
    # how many rows each neighbour can provide / we can provide: bands may be shorter than radius
    mine = torch.tensor([rows], dtype=torch.int64, device=band.device)
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    sizes = [int(s.item()) for s in sizes]
    top = min(radius, sizes[rank - 1]) if rank > 0 else 0
    bottom = min(radius, sizes[rank + 1]) if rank < world - 1 else 0
    check_halo_reach(sizes, radius)
    ext = torch.empty((top + rows + bottom,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
    ext[top:top + rows].copy_(band)
    ops = []
    send_up = min(radius, rows)
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, band[:send_up].contiguous(), rank - 1, group=group))
        ops.append(dist.P2POp(dist.irecv, ext[:top], rank - 1, group=group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, band[rows - send_up:].contiguous(), rank + 1, group=group))
        ops.append(dist.P2POp(dist.irecv, ext[top + rows:], rank + 1, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return ext, top, bottom


def blur_band(band, radius: int, lut: np.ndarray, oob_rgbx: int, rank: int, world: int, group=None):
    """Blur one row band of a canvas split across `world` GPUs, in place.  band: CUDA uint8 tensor
    [rows, width, 4].  Rows at the true image border see `oob_rgbx`; interior cuts see the
    neighbours' rows."""
    import torch

    from . import device as dev

    dev.set_stream(torch_stream_handle())  # NCCL and the kernels share one stream
    try:
        ext, top, bottom = exchange_halos(band, radius, rank, world, group)
        rows, width = band.shape[0], band.shape[1]
        img = dev.DeviceImage.wrap(ext.data_ptr(), width, ext.shape[0], owner=ext)
        dev.blur_rows(img, lut, radius, oob_rgbx, top, top + rows)
        band.copy_(ext[top:top + rows])
    finally:
        dev.set_stream(None)  # back to the library's stream (ordered after what was queued here)
    return band


class RowBand:
    """This rank's row band of one canvas split across `world` ranks, stored with `margin` spare rows above and
    below it: the neighbours' halo rows are received straight into the margins and the row-band kernel runs on
    [halo ; band ; halo] in place — no size exchange (band heights follow from `band_range`), no staging copies.

    rb = RowBand(height, width, rank, world, margin=64); rb.band[...] = pixels; rb.blur(32, lut, 0)"""

    def __init__(self, height: int, width: int, rank: int, world: int, margin: int, device="cuda", group=None):
        import torch

        self.height, self.width, self.rank, self.world, self.margin, self.group = height, width, rank, world, margin, group
        self.sizes = [band_range(height, world, k)[1] - band_range(height, world, k)[0] for k in range(world)]
        self.rows = self.sizes[rank]
        self.buf = torch.zeros((margin + self.rows + margin, width, 4), dtype=torch.uint8, device=device)
        self.band = self.buf[margin:margin + self.rows]

    def halo_rows(self, radius: int):
        """(top, bottom): halo rows this rank receives for a filter of `radius` rows."""
        r, w, sizes = self.rank, self.world, self.sizes
        if radius > self.margin:
            raise ValueError("radius exceeds the band's margin")
        check_halo_reach(sizes, radius)
        top = min(radius, sizes[r - 1]) if r > 0 else 0
        bottom = min(radius, sizes[r + 1]) if r < w - 1 else 0
        return top, bottom

    def exchange(self, radius: int):
        """Fill the margins with the neighbours' rows (point-to-point send / recv).  Returns (ext, top, bottom) with
        ext = the contiguous view [halo_top ; band ; halo_bottom] of the buffer."""
        import torch.distributed as dist

        ops, top, bottom = self._halo_ops(radius)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        m = self.margin
        return self.buf[m - top:m + self.rows + bottom], top, bottom

    def _halo_ops(self, radius: int):
        import torch.distributed as dist

        top, bottom = self.halo_rows(radius)
        m, rows, r, w = self.margin, self.rows, self.rank, self.world
        send = min(radius, rows)
        ops = []
        if r > 0:
            ops.append(dist.P2POp(dist.isend, self.band[:send], r - 1, group=self.group))
            ops.append(dist.P2POp(dist.irecv, self.buf[m - top:m], r - 1, group=self.group))
        if r < w - 1:
            ops.append(dist.P2POp(dist.isend, self.band[rows - send:], r + 1, group=self.group))
            ops.append(dist.P2POp(dist.irecv, self.buf[m + rows:m + rows + bottom], r + 1, group=self.group))
        return ops, top, bottom

    def _ext_image(self, top: int, bottom: int):
        """DeviceImage over [halo_top ; band ; halo_bottom] of the buffer (cached per halo size)."""
        from . import device as dev

        key = (top, bottom)
        cache = self.__dict__.setdefault("_ext_cache", {})
        if key not in cache:
            ext = self.buf[self.margin - top:self.margin + self.rows + bottom]
            cache[key] = dev.DeviceImage.wrap(ext.data_ptr(), self.width, ext.shape[0], owner=self.buf)
        return cache[key]

    def blur(self, radius: int, lut: np.ndarray, oob_rgbx: int, overlap: bool = True):
        """Blur the band in place as part of the whole canvas: rows at the true image border see `oob_rgbx`, interior
        cuts see the neighbours' rows.

        overlap=True: the halo exchange runs on a side stream WHILE the horizontal pass — which
        needs no halo — blurs the band's own rows on the main stream; when the neighbours' rows have arrived their
        2 x radius rows get the horizontal pass and the vertical pass finishes the band.  NCCL's fixed send / recv
        latency (~0.2 ms for 4 MiB that NVLink moves in ~5 us) is hidden behind the X pass instead of preceding it."""
        import torch

        from . import device as dev

        main = torch.cuda.current_stream()
        dev.set_stream(torch_stream_handle())  # the kernels are ordered with torch's work on this stream
        try:
            if not overlap or radius > 64:
                ext, top, bottom = self.exchange(radius)
                dev.blur_rows(self._ext_image(top, bottom), lut, radius, oob_rgbx, top, top + self.rows)
                return self.band
            import torch.distributed as dist

            ops, top, bottom = self._halo_ops(radius)
            img = self._ext_image(top, bottom)
            side = self.__dict__.get("_side")
            if side is None:
                side = self._side = torch.cuda.Stream()
            if ops:
                side.wait_stream(main)  # the band's pixels are complete before they are sent
                with torch.cuda.stream(side):
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()  # stream-side wait: `side` continues when the halo rows are in place
            dev.blur_rows_x(img, lut, radius, oob_rgbx, top, top + self.rows)  # own rows: no halo needed
            if ops:
                main.wait_stream(side)
            if top:
                dev.blur_rows_x(img, lut, radius, oob_rgbx, 0, top)
            if bottom:
                dev.blur_rows_x(img, lut, radius, oob_rgbx, top + self.rows, top + self.rows + bottom)
            dev.blur_rows_y(img, lut, radius, oob_rgbx, top, top + self.rows)
            return self.band
        finally:
            dev.set_stream(None)

    def spread(self, amount: int):
        """spread (images.nim:700-758) of the band as part of the whole canvas: |amount| halo rows per cut."""
        from . import device as dev

        dev.set_stream(torch_stream_handle())
        try:
            ext, top, bottom = self.exchange(abs(amount))
            dev.spread_rows(self._ext_image(top, bottom), amount, top, top + self.rows)
            return self.band
        finally:
            dev.set_stream(None)

    def shadow(self, offset, spread: int, radius: int, lut: np.ndarray, rgbx: int):
        """shadow (images.nim:760-776) of the band as part of the whole canvas.  Returns a tensor [rows, width, 4] with
        the band's rows of the shadow image; the band itself is left unchanged.  Halo = ceil|offset.y| + |spread| +
        radius input rows per cut (offset copy, spread and blur each reach that far)."""
        import math

        import torch

        from . import device as dev

        need = int(math.ceil(abs(offset[1]))) + abs(spread) + max(radius, 0)
        dev.set_stream(torch_stream_handle())
        try:
            ext, top, bottom = self.exchange(need)
            out = self.__dict__.get("_shadow_out")
            if out is None or out.shape[0] != ext.shape[0]:
                out = self._shadow_out = torch.empty_like(ext)
            dst = dev.DeviceImage.wrap(out.data_ptr(), self.width, out.shape[0], owner=out)
            dev.shadow_rows(self._ext_image(top, bottom), dst, float(offset[0]), float(offset[1]), spread, lut, radius, rgbx,
                            top, top + self.rows)
            return out[top:top + self.rows]
        finally:
            dev.set_stream(None)

    def blend(self, src_band, mode: int):
        """dst.draw(src, blendMode) restricted to this band (images.nim:468-529): per-pixel, no exchange."""
        from . import device as dev

        dev.set_stream(torch_stream_handle())
        try:
            d = dev.DeviceImage.wrap(self.band.data_ptr(), self.width, self.rows, owner=self.buf)
            s_ = dev.DeviceImage.wrap(src_band.data_ptr(), self.width, self.rows, owner=src_band)
            dev.blend_rect(d, s_, 0, 0, mode)
            return self.band
        finally:
            dev.set_stream(None)

    def fill(self, cmdlist, count_covered=False):
        """Rasterise this band's rows of a whole-canvas command list (fills: scanlines are independent given the
        segment list; partition boundaries stay those of the whole canvas — pixie_cuda_cmdlist_run_rows)."""
        from . import device as dev

        y0, _ = band_range(self.height, self.world, self.rank)
        dev.set_stream(torch_stream_handle())
        try:
            d = dev.DeviceImage.wrap(self.band.data_ptr(), self.width, self.rows, owner=self.buf)
            return cmdlist.run_rows(d, y0, y0 + self.rows, count_covered)
        finally:
            dev.set_stream(None)
