"""N>1 host logic on CPU: world_size-2 gloo run of the row-band halo exchange used between the reuse stages."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, halo, out, balanced=False, deferred=False):
    sys.path.insert(0, ROOT)
    from volumetricrestirrelease_b200.multi_gpu import balanced_row_bands, exchange_row_halo, row_bands
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if balanced:   # cost concentrated in the lower rows -> uneven bands, every rank derives the same cut
        cost = np.where(np.arange(H) >= H // 2, 10.0, 1.0) * W
        bands = balanced_row_bands(cost, world, min_rows=16)
        assert bands[0][1] - bands[0][0] != bands[1][1] - bands[1][0]
    else:
        bands = row_bands(H, world)
    r0, r1 = bands[rank]
    full = torch.arange(H * W * 16, dtype=torch.int64).remainder(251).to(torch.uint8).view(H, W * 16)
    planes = []
    for k in range(2):                       # two SoA planes like a reservoir buffer
        t = torch.zeros_like(full)
        t[r0:r1] = full[r0:r1] + k
        planes.append(t)
    works = exchange_row_halo(planes, (r0, r1), halo, rank, world, wait=not deferred)
    for w in works:          # deferred mode: the caller overlaps the transfer with other work and waits later
        w.wait()
    ok = True
    lo, hi = max(0, r0 - halo), min(H, r1 + halo)
    for k, t in enumerate(planes):
        ok &= bool(torch.equal(t[lo:hi], (full[lo:hi] + k)))
        if lo > 0:
            ok &= bool((t[:lo] == 0).all())          # nothing beyond the halo is touched
        if hi < H:
            ok &= bool((t[hi:] == 0).all())
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_world_size_2():
    world, H, W, halo = 2, 48, 20, 10
    port = _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, H, W, halo, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_halo_exchange_balanced_bands_deferred_wait():
    """Uneven (cost-balanced) bands and the deferred-wait form used for the history halo."""
    world, H, W, halo = 2, 64, 12, 10
    port = _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, H, W, halo, out, True, True), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
