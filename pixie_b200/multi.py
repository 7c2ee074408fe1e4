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
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        cpus, node = set(), None
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(handle).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bus = bus.lower()
            if len(bus.split(":")[0]) == 8:  # nvml reports an 8-digit domain, sysfs uses 4
                bus = bus[4:]
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
            if node >= 0:
                with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                    for part in f.read().strip().split(","):
                        a, _, b = part.partition("-")
                        cpus.update(range(int(a), int(b or a) + 1))
        except Exception:
            node = None
        if not cpus:
            # virtualised hosts report numa_node = -1: ask the driver for the GPU's ideal CPU set instead
            # (the "CPU Affinity" column of `nvidia-smi topo -m`)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
            for wi, word in enumerate(mask):
                for bit in range(64):
                    if (int(word) >> bit) & 1:
                        cpus.add(wi * 64 + bit)
            info["source"] = "nvmlDeviceGetCpuAffinity"
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update({"numa_node": node if node is not None and node >= 0 else None, "cpus": len(allowed)})
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

    def __init__(self, height: int, width: int, rank: int, world: int, margin: int, device="cuda", group=None,
                 transport: str = "auto"):
        """transport: how halo rows travel.  "nccl": torch.distributed point-to-point send / recv (NCCL over NVLink on
        GPUs, gloo in the CPU tests).  "peer": the band buffers are CUDA-IPC mapped into the neighbours' processes and a
        small kernel stores the edge rows straight into the neighbour's margin over NVLink, followed by an epoch flag
        (pixie_cuda_halo_push / _wait) — no collective launch, a few microseconds instead of a send / recv pair's
        ~0.1 ms.  "auto": "peer" on CUDA devices when the IPC mapping works, else "nccl"."""
        import torch

        self.height, self.width, self.rank, self.world, self.margin, self.group = height, width, rank, world, margin, group
        self.sm_reserve = 0
        self.sizes = [band_range(height, world, k)[1] - band_range(height, world, k)[0] for k in range(world)]
        self.rows = self.sizes[rank]
        self.cur = 0      # which of the two buffers holds the band (blur is out of place and swaps them)
        self.epoch = 0    # halo exchanges so far (identical on every rank: exchanges are collective calls)
        self.transport = "nccl"
        self._peer = None
        on_cuda = str(device).startswith("cuda")
        if transport in ("auto", "peer") and on_cuda and world > 1:
            try:
                self._init_peer()
                self.transport = "peer"
            except Exception as e:  # no IPC in this container, no peer access: fall back to send / recv
                if transport == "peer":
                    raise
                self.peer_error = repr(e)[:120]
        if self.transport == "nccl":
            self.buf = torch.zeros((margin + self.rows + margin, width, 4), dtype=torch.uint8, device=device)
            self.buf2 = None
        self.band = self.buf[margin:margin + self.rows]

    # ---- peer transport: one IPC allocation per rank = [flags 256 B][buffer 0][buffer 1]
    def _buf_bytes(self, k: int) -> int:
        return ((self.margin + self.sizes[k] + self.margin) * self.width * 4 + 255) & ~255

    def _init_peer(self):
        import torch
        import torch.distributed as dist

        from . import device as dev

        nbytes = 256 + 2 * self._buf_bytes(self.rank)
        mine = dev.PeerBuffer(nbytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, mine.handle, group=self.group)
        peers = {}
        for k in (self.rank - 1, self.rank + 1):
            if 0 <= k < self.world:
                peers[k] = dev.PeerBuffer.open(handles[k], 256 + 2 * self._buf_bytes(k))
        dist.barrier(group=self.group)  # every mapping exists before anybody pushes
        self._peer = {"mine": mine, "peers": peers}
        flat = torch.as_tensor(mine, device="cuda")
        bb = self._buf_bytes(self.rank)
        shape = (self.margin + self.rows + self.margin, self.width, 4)
        nb = shape[0] * shape[1] * 4
        self.buf = flat[256:256 + nb].view(shape)
        self.buf2 = flat[256 + bb:256 + bb + nb].view(shape)
        self._flat = flat

    def _peer_exchange_start(self, radius: int):
        """Stream-ordered: tell the neighbours this rank's margins are free, wait for theirs, push the edge rows."""
        from . import device as dev

        self.epoch += 1
        e, r, w, m = self.epoch, self.rank, self.world, self.margin
        top, bottom = self.halo_rows(radius)
        send = min(radius, self.rows)
        row_bytes = self.width * 4
        mine, peers = self._peer["mine"], self._peer["peers"]
        # flags of a rank (uint32, written by its neighbours): [0] rows from above arrived, [1] rows from below arrived,
        # [2] the upper neighbour's bottom margin is free, [3] the lower neighbour's top margin is free
        up = down = None
        if r > 0:
            k = r - 1
            up = dev.HaloDir(self.band.data_ptr(), peers[k].ptr + 256 + self.cur * self._buf_bytes(k) + (m + self.sizes[k]) * row_bytes,
                             send * row_bytes, peers[k].ptr + 12, peers[k].ptr + 4, mine.ptr + 8)
        if r < w - 1:
            k = r + 1
            down = dev.HaloDir(self.band[self.rows - send:].data_ptr(), peers[k].ptr + 256 + self.cur * self._buf_bytes(k) + (m - send) * row_bytes,
                               send * row_bytes, peers[k].ptr + 8, peers[k].ptr + 0, mine.ptr + 12)
        dev.halo_exchange(up, down, e)
        return top, bottom

    def _peer_exchange_finish(self):
        from . import device as dev

        mine = self._peer["mine"]
        dev.halo_wait2(mine.ptr + 0 if self.rank > 0 else None, mine.ptr + 4 if self.rank < self.world - 1 else None, self.epoch)

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

        m = self.margin
        if self.transport == "peer":
            from . import device as dev

            dev.set_stream(torch_stream_handle())
            top, bottom = self._peer_exchange_start(radius)
            self._peer_exchange_finish()
            return self.buf[m - top:m + self.rows + bottom], top, bottom
        ops, top, bottom = self._halo_ops(radius)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
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

    def _ext_image(self, top: int, bottom: int, buf=None):
        """DeviceImage over [halo_top ; band ; halo_bottom] of a buffer (cached per buffer and halo size)."""
        from . import device as dev

        buf = self.buf if buf is None else buf
        key = (buf.data_ptr(), top, bottom)
        cache = self.__dict__.setdefault("_ext_cache", {})
        if key not in cache:
            ext = buf[self.margin - top:self.margin + self.rows + bottom]
            cache[key] = dev.DeviceImage.wrap(ext.data_ptr(), self.width, ext.shape[0], owner=buf)
        return cache[key]

    def blur(self, radius: int, lut: np.ndarray, oob_rgbx: int, overlap: bool = True):
        """Blur the band as part of the whole canvas: rows at the true image border see `oob_rgbx`, interior cuts see
        the neighbours' rows.  Returns the band (``self.band``), which afterwards lives in the band's SECOND buffer:
        the blur is out of place (pixie_cuda_blur_rows_to), the two buffers swap roles on every call.

        overlap=True: the band's interior — which depends on the band's own rows only — is blurred
        while the halo rows travel.  transport "peer": one exchange kernel (stores into the neighbours' margins + epoch
        flags) and ONE blur kernel behind it whose tiles wait for the flags only if they read halo rows.  transport
        "nccl": the send / recv pair is started first, the interior rows [radius, rows - radius) are blurred, and
        after `wait()` the two edge strips of `radius` rows follow."""
        import torch

        from . import device as dev

        main = torch.cuda.current_stream()
        if self.buf2 is None:
            self.buf2 = torch.zeros_like(self.buf)
        dev.set_stream(torch_stream_handle())  # the kernels are ordered with torch's work on this stream
        try:
            import torch.distributed as dist

            peer = self.transport == "peer"
            if peer:
                top, bottom = self._peer_exchange_start(radius)  # the rows are on their way over NVLink
                ops = []
            else:
                ops, top, bottom = self._halo_ops(radius)
            src, dst = self._ext_image(top, bottom), self._ext_image(top, bottom, self.buf2)
            b0, b1 = top, top + self.rows
            split = overlap and (peer or ops) and self.rows > 2 * radius and self.world > 1
            if peer:
                # ONE blur launch right behind the exchange kernel: its interior tiles start at once, the tiles that
                # read halo rows wait — inside the kernel — for the epoch flag the neighbour publishes after its rows
                mine = self._peer["mine"]
                if overlap:
                    dev.blur_rows_to_flags(src, dst, lut, radius, oob_rgbx, b0, b1, mine.ptr + 0 if self.rank > 0 else None,
                                           mine.ptr + 4 if self.rank < self.world - 1 else None, self.epoch)
                else:
                    self._peer_exchange_finish()
                    dev.blur_rows_to(src, dst, lut, radius, oob_rgbx, b0, b1)
            elif split:
                # torch's NCCL process group runs send / recv on its own stream, ordered after the work already queued
                # on this one; `req.wait()` makes this stream wait for them — so the exchange is simply started first
                # and waited for after the interior rows
                reqs = dist.batch_isend_irecv(ops)
                dev.set_sm_reserve(self.sm_reserve)  # optionally leave SMs to NCCL beside the one-CTA-per-SM blur
                dev.blur_rows_to(src, dst, lut, radius, oob_rgbx, b0 + radius, b1 - radius)  # no halo row is read
                dev.set_sm_reserve(0)
                for req in reqs:
                    req.wait()
                dev.blur_rows_to(src, dst, lut, radius, oob_rgbx, b0, b0 + radius)
                dev.blur_rows_to(src, dst, lut, radius, oob_rgbx, b1 - radius, b1)
            else:
                if ops:
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
                dev.blur_rows_to(src, dst, lut, radius, oob_rgbx, b0, b1)
            self.cur ^= 1
            self.buf, self.buf2 = self.buf2, self.buf
            self.band = self.buf[self.margin:self.margin + self.rows]
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
