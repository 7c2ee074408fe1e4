"""Pieces of the row-band blur at N GPUs (torchrun): exchange alone, blur pieces alone, overlapped and serial totals."""
import os
import statistics
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixie_b200 import device as dev, host, multi, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev.init(local)
n, r = 16384, 32
lut = host.gaussianKernel(r)
transport = sys.argv[1] if len(sys.argv) > 1 else "auto"
rb = multi.RowBand(n, n, rank, world, margin=64, transport=transport)
rb.band.copy_(torch.from_numpy(synth.random_premultiplied(256, n, rank)).cuda().repeat((rb.rows + 255) // 256, 1, 1)[: rb.rows])
if rb.buf2 is None:
    rb.buf2 = torch.zeros_like(rb.buf)


def timed(fn, reps=6):
    ts = []
    for it in range(reps):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it > 1:
            ts.append(e0.elapsed_time(e1))
    t = torch.tensor([statistics.median(ts)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


dev.set_stream(multi.torch_stream_handle())
top, bottom = rb.halo_rows(r)
src, dst = rb._ext_image(top, bottom), rb._ext_image(top, bottom, rb.buf2)
b0, b1 = top, top + rb.rows
res = {}
res["exchange"] = timed(lambda: rb.exchange(r))
res["blur whole band"] = timed(lambda: dev.blur_rows_to(src, dst, lut, r, 0, b0, b1))
res["blur interior"] = timed(lambda: dev.blur_rows_to(src, dst, lut, r, 0, b0 + r, b1 - r))
res["blur 2 edges"] = timed(lambda: (dev.blur_rows_to(src, dst, lut, r, 0, b0, b0 + r), dev.blur_rows_to(src, dst, lut, r, 0, b1 - r, b1)))
res["blur 1 edge"] = timed(lambda: dev.blur_rows_to(src, dst, lut, r, 0, b0, b0 + r))
dev.set_stream(None)
res["overlapped"] = timed(lambda: rb.blur(r, lut, 0, overlap=True))
res["serial"] = timed(lambda: rb.blur(r, lut, 0, overlap=False))
if rank == 0:
    print(rb.transport, getattr(rb, "peer_error", ""), {k: round(v, 3) for k, v in res.items()})
dist.destroy_process_group()
