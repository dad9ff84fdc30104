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


def _gather_worker(rank, world, port, H, W, out):
    sys.path.insert(0, ROOT)
    from volumetricrestirrelease_b200.multi_gpu import gather_row_bands
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bands = [(0, 24), (24, H)]                 # uneven bands
    r0, r1 = bands[rank]
    full = torch.arange(H * W * 8, dtype=torch.int64).remainder(239).to(torch.uint8).view(H, W * 8)
    planes = []
    for k in range(3):                         # two reservoir planes + features
        t = torch.zeros_like(full)
        t[r0:r1] = full[r0:r1] + k
        planes.append(t)
    for w in gather_row_bands(planes, bands, rank):
        w.wait()
    out[rank] = int(all(torch.equal(t, full + k) for k, t in enumerate(planes)))
    dist.barrier()
    dist.destroy_process_group()


def test_history_all_gather_fallback_world_size_2():
    """The temporal-history fallback for motion beyond the halo: after gather_row_bands every rank holds every band."""
    world, H, W = 2, 56, 10
    port = _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_gather_worker, args=(world, port, H, W, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_reprojection_row_bound():
    """The host-side motion bound that decides between the history halo and the all-gather: zero motion stays inside the margin,
    the bound grows with the camera step, and it really bounds the row displacement of points inside the box."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import env_scene
    from volumetricrestirrelease_b200.multi_gpu import reprojection_row_bound
    sc = env_scene()
    lo, hi = sc.volume_bounds_world()
    corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    W, H = 640, 360
    c0 = sc.camera.data(W, H)
    assert reprojection_row_bound(c0, c0, corners, H) == 2.0
    p0 = np.array(sc.camera.position)
    rng = np.random.default_rng(3)
    pts = lo + rng.random((4000, 3)) * (hi - lo)
    last = 0.0
    for step in (0.5, 3.0, 12.0):
        sc.camera.position = tuple(p0 + np.array([0.3, 1.0, 0.1]) * step)
        c1 = sc.camera.data(W, H)
        b = reprojection_row_bound(c0, c1, corners, H)
        assert b > last
        last = b

        def rows(cam):
            V = np.array(list(cam.viewMat), dtype=np.float64).reshape(4, 4)
            P = np.array(list(cam.projMat), dtype=np.float64).reshape(4, 4)
            c = np.concatenate([pts, np.ones((len(pts), 1))], axis=1) @ V @ P
            return (-0.5 * c[:, 1] / c[:, 3] + 0.5) * H
        assert np.abs(rows(c0) - rows(c1)).max() <= b
