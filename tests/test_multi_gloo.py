"""The N>1 host logic on CPU: world_size 2 and 3 over gloo (no GPU): shard ranges, the halo
exchange, and that band + halo blurs reproduce the global blur (checked with the oracle)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, h, w, radius, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from pixie_b200 import host, multi, synth
    from _oracle import OracleBackend

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = synth.random_premultiplied(h, w, 123)  # every rank builds the same global image
        y0, y1 = multi.band_range(h, world, rank)
        band = torch.from_numpy(np.ascontiguousarray(img[y0:y1]))
        ext, top, bottom = multi.exchange_halos(band, radius, rank, world)
        e0, e1 = y0 - top, y1 + bottom
        assert np.array_equal(ext.numpy(), img[e0:e1]), "halo rows are not the neighbours' rows"
        assert top == (min(radius, y0) if rank > 0 else 0)
        # band + halo blurred with the oracle == the rows of the global blur
        lut = host.gaussianKernel(radius)
        ob = OracleBackend(0)
        whole = img.copy()
        ob.blur(whole, lut, radius, 0)
        part = np.ascontiguousarray(ext.numpy().copy())
        ob.blur(part, lut, radius, 0)
        assert np.array_equal(part[top:top + (y1 - y0)], whole[y0:y1]), "banded blur differs from the global blur"
        # the same exchange with the band stored between its halo margins (multi.RowBand): no staging copies
        rb = multi.RowBand(h, w, rank, world, margin=radius + 3, device="cpu")
        rb.band.copy_(band)
        ext2, top2, bottom2 = rb.exchange(radius)
        assert (top2, bottom2) == (top, bottom) and ext2.is_contiguous()
        assert np.array_equal(ext2.numpy(), img[e0:e1]), "RowBand halo rows are not the neighbours' rows"
        assert np.array_equal(rb.band.numpy(), img[y0:y1])
        # shadow / spread bands: halo = ceil|offset.y| + |spread| + radius input rows; band + halo run
        # through the oracle's shadow / spread == the rows of the whole-image result
        off, sp, sr = (3.0, -4.0), 2, 5
        need = 4 + sp + sr
        rb2 = multi.RowBand(h, w, rank, world, margin=need, device="cpu")
        rb2.band.copy_(band)
        ext3, top3, bottom3 = rb2.exchange(need)
        slut = host.gaussianKernel(sr)
        whole_sh = ob.shadow(img, off[0], off[1], sp, slut, sr, 0xC8000000)
        part_sh = ob.shadow(np.ascontiguousarray(ext3.numpy().copy()), off[0], off[1], sp, slut, sr, 0xC8000000)
        assert np.array_equal(part_sh[top3:top3 + (y1 - y0)], whole_sh[y0:y1]), "banded shadow differs from the global shadow"
        whole_sp = img.copy()
        ob.spread(whole_sp, -sp)
        ext4, top4, _ = rb2.exchange(sp)
        part_sp = np.ascontiguousarray(ext4.numpy().copy())
        ob.spread(part_sp, -sp)
        assert np.array_equal(part_sp[top4:top4 + (y1 - y0)], whole_sp[y0:y1]), "banded spread differs from the global spread"
        # the multi-hop check is a function of the (rank-independent) band sizes: every rank raises, or none does
        try:
            multi.check_halo_reach([33, 33, 32, 32], 33)
            raised = False
        except ValueError:
            raised = True
        assert raised
        multi.check_halo_reach([20, 40, 40, 10], 33)  # short EDGE bands are fine: beyond them is the image border
        # shards of independent units cover the range exactly once
        ranges = [multi.shard_range(1001, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 1001 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_banded_blur_halo_exchange(world, tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + world + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, 150, 64, 12, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
