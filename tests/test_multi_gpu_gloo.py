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


# ---- the orchestration of ShardedPass.execute against a model pass -----------------------------------------------------------
class _ModelPass:
    """Stands in for the CUDA pass on CPU tensors: every stage is a deterministic row function that READS what the real stage
    reads (K2: the history around the row; K3: the round's input buffer within the sample radius, with its extra-bounce records and
    p_partial plane) and writes only its own band — so a missing or misplaced exchange changes the band's result."""

    def __init__(self, H, W, band, params, halo_t):
        from volumetricrestirrelease_b200 import capi
        self.capi, self.H, self.W, self.band, self.params, self.halo_t = capi, H, W, band, params, halo_t
        widths = {capi.BUF_RESERVOIR_0: 32, capi.BUF_RESERVOIR_1: 32, capi.BUF_RESERVOIR_TEMPORAL: 32, capi.BUF_FEATURES: 8, capi.BUF_FEATURES_TEMPORAL: 8,
                  capi.BUF_EXTRA_0: 12 * (params.mMaxBounces - 1), capi.BUF_EXTRA_1: 12 * (params.mMaxBounces - 1),
                  capi.BUF_EXTRA_TEMPORAL: 12 * (params.mMaxBounces - 1), capi.BUF_PPARTIAL_0: 4, capi.BUF_PPARTIAL_1: 4, capi.BUF_PPARTIAL_TEMPORAL: 4}
        self.buf = {b: torch.zeros(H, max(w, 1) * W, dtype=torch.int64) for b, w in widths.items()}
        self.frame = 0
        self.final = None

    def frame_count(self):
        return self.frame

    def spatial_input_buffer(self, r):
        return self.capi.BUF_RESERVOIR_0 if r % 2 == 0 else self.capi.BUF_RESERVOIR_1

    def _rows(self):
        return range(self.band[0], self.band[1])

    def _around(self, t, y, h):
        return int(t[max(0, y - h):min(self.H, y + h + 1)].sum())

    def execute_stage(self, stage, arg=0, *_):
        c, b, B = self.capi, self.buf, self.params.mMaxBounces
        vr = bool(self.params.mVertexReuse) and B > 1
        twin = {c.BUF_RESERVOIR_0: (c.BUF_EXTRA_0, c.BUF_PPARTIAL_0), c.BUF_RESERVOIR_1: (c.BUF_EXTRA_1, c.BUF_PPARTIAL_1),
                c.BUF_RESERVOIR_TEMPORAL: (c.BUF_EXTRA_TEMPORAL, c.BUF_PPARTIAL_TEMPORAL)}

        def family(res):
            return [res] + ([twin[res][0]] if B > 1 else []) + ([twin[res][1]] if vr else [])
        if stage == 0:
            for y in self._rows():
                b[c.BUF_FEATURES][y] = 7 * y + self.frame
        elif stage == 1:
            for k, t in enumerate(family(c.BUF_RESERVOIR_0)):
                for y in self._rows():
                    t = b[family(c.BUF_RESERVOIR_0)[k]]
                    t[y] = 1000 * (k + 1) + 13 * y + self.frame
        elif stage == 2 and self.frame > 0:
            new = {}
            for cur, prv in zip(family(c.BUF_RESERVOIR_0), family(c.BUF_RESERVOIR_TEMPORAL)):
                new[cur] = {y: int(b[cur][y, 0]) + self._around(b[prv], y, self.halo_t) + self._around(b[c.BUF_FEATURES_TEMPORAL], y, self.halo_t) for y in self._rows()}
            for cur, rows in new.items():
                for y, v in rows.items():
                    b[cur][y] = v % 1000003
        elif stage == 3:
            src = self.spatial_input_buffer(arg)
            dst = c.BUF_RESERVOIR_1 if src == c.BUF_RESERVOIR_0 else c.BUF_RESERVOIR_0
            h = int(np.ceil(self.params.mSampleRadius))
            for s, d in zip(family(src), family(dst)):
                for y in self._rows():
                    b[d][y] = (self._around(b[s], y, h) + int(b[c.BUF_FEATURES][y, 0])) % 1000003     # features: the own row only
            self.final = dst
        elif stage == 4:
            src = self.final if self.final is not None else c.BUF_RESERVOIR_0
            for s, d in zip(family(src), family(c.BUF_RESERVOIR_TEMPORAL)):
                for y in self._rows():
                    b[d][y] = b[s][y]
            for y in self._rows():
                b[c.BUF_FEATURES_TEMPORAL][y] = b[c.BUF_FEATURES][y]
        elif stage == 6:
            self.frame += 1


def _orchestration_worker(rank, world, port, H, W, vertex_reuse, out):
    sys.path.insert(0, ROOT)
    from volumetricrestirrelease_b200 import VolumetricReSTIRParams, capi
    from volumetricrestirrelease_b200.multi_gpu import ShardedPass, row_bands
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = VolumetricReSTIRParams(mMaxBounces=3, mVertexReuse=vertex_reuse, mSpatialReuseRounds=2, mSampleRadius=5.0)

    class Model(ShardedPass):
        def _planes(self, buffer):
            return [self.p.buf[buffer]]

        def _history_needs_gather(self):
            return False

    def run(world_, rank_):
        band = row_bands(H, world_)[rank_]
        mp_ = _ModelPass(H, W, band, params, halo_t=6)
        mp_._scene = type("S", (), {"camera": type("C", (), {"data": staticmethod(lambda w, h: None)})()})()
        sp = Model(mp_, W, H, rank_, world_, torch.device("cpu"), temporal_halo=6)
        for _ in range(3):
            if world_ == 1:          # the whole-frame call of the real pass: the same stages, no exchange
                for st, arg in [(0, 0), (1, 0), (2, 0), (3, 0), (3, 1), (4, 0), (5, 0), (6, 0)]:
                    mp_.execute_stage(st, arg)
            else:
                sp.execute(0)
        for w in getattr(sp, "_pending_history", []):
            w.wait()
        return mp_, band

    sharded, band = run(world, rank)
    whole, _ = run(1, 0)
    ok = True
    for bid in (capi.BUF_RESERVOIR_TEMPORAL, capi.BUF_EXTRA_TEMPORAL, capi.BUF_FEATURES_TEMPORAL) + ((capi.BUF_PPARTIAL_TEMPORAL,) if vertex_reuse else ()):
        ok &= bool(torch.equal(sharded.buf[bid][band[0]:band[1]], whole.buf[bid][band[0]:band[1]]))
    if not vertex_reuse:             # the p_partial plane does not travel without vertex reuse
        other = row_bands(H, world)[1 - rank]
        ok &= bool((sharded.buf[capi.BUF_PPARTIAL_TEMPORAL][other[0]:other[1]] == 0).all())
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_execute_exchanges_what_every_stage_reads():
    """ShardedPass.execute on a model pass (world 2 vs the un-sharded frame): reservoir planes, extra-bounce records and — with
    vertex reuse — the p_partial plane reach the neighbour before each spatial round, and the history before the next frame's K2."""
    for vertex_reuse in (0, 1):
        world, H, W = 2, 48, 3
        port = _free_port()
        out = mp.get_context("spawn").Manager().dict()
        mp.spawn(_orchestration_worker, args=(world, port, H, W, vertex_reuse, out), nprocs=world, join=True)
        assert dict(out) == {0: 1, 1: 1}, vertex_reuse
