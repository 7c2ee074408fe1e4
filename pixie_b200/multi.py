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
    if (rank > 0 and sizes[rank - 1] < radius and rank - 1 > 0) or \
       (rank < world - 1 and sizes[rank + 1] < radius and rank + 1 < world - 1):
        raise ValueError("bands shorter than the blur radius need multi-hop halos: use fewer ranks")
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
    ext, top, bottom = exchange_halos(band, radius, rank, world, group)
    rows, width = band.shape[0], band.shape[1]
    img = dev.DeviceImage.wrap(ext.data_ptr(), width, ext.shape[0], owner=ext)
    dev.blur_rows(img, lut, radius, oob_rgbx, top, top + rows)
    band.copy_(ext[top:top + rows])
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
        if (r > 0 and sizes[r - 1] < radius and r - 1 > 0) or (r < w - 1 and sizes[r + 1] < radius and r + 1 < w - 1):
            raise ValueError("bands shorter than the blur radius need multi-hop halos: use fewer ranks")
        top = min(radius, sizes[r - 1]) if r > 0 else 0
        bottom = min(radius, sizes[r + 1]) if r < w - 1 else 0
        return top, bottom

    def exchange(self, radius: int):
        """Fill the margins with the neighbours' rows (point-to-point send / recv).  Returns (ext, top, bottom) with
        ext = the contiguous view [halo_top ; band ; halo_bottom] of the buffer."""
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
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return self.buf[m - top:m + rows + bottom], top, bottom

    def blur(self, radius: int, lut: np.ndarray, oob_rgbx: int):
        """Blur the band in place as part of the whole canvas: rows at the true image border see `oob_rgbx`, interior
        cuts see the neighbours' rows."""
        import torch

        from . import device as dev

        dev.set_stream(torch_stream_handle())  # NCCL and the kernels share one stream
        ext, top, bottom = self.exchange(radius)
        img = dev.DeviceImage.wrap(ext.data_ptr(), self.width, ext.shape[0], owner=self.buf)
        dev.blur_rows(img, lut, radius, oob_rgbx, top, top + self.rows)
        return self.band
