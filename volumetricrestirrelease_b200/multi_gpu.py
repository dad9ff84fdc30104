"""Row-sharded multi-GPU execution of the pass: one process per GPU, contiguous row bands, reservoir-halo exchange.

The path shards by image rows: K0/K1/K5 are per-pixel independent; K3 reads neighbour reservoirs
within ``mSampleRadius`` rows of the round's *input* buffer and K2 reads the previous frame's reservoir/features at the
reprojected pixel.  So the only data that crosses GPUs is a halo of reservoir rows (and feature rows for K2), exchanged
between neighbouring ranks with NCCL send/recv directly on the pass's device buffers (zero-copy tensor views).
The volume, env map and lights are replicated.  RNG is keyed by absolute pixel + frame, so a sharded frame is
bit-identical to the single-GPU frame as long as the halo covers the taps (tests/test_multi_gpu.py).

The same code runs on CPU tensors with the gloo backend (host-logic tests, world_size 2).
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _capi as capi


def row_bands(height, world_size, align=8):
    """Contiguous bands, `align`-row aligned (CTA tile height), as even as possible.  Returns [(r0, r1)] per rank."""
    units = math.ceil(height / align)
    base, rem = divmod(units, world_size)
    bands, u = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        bands.append((min(height, u * align), min(height, (u + n) * align)))
        u += n
    return bands


def balanced_row_bands(row_cost, world_size, align=8, min_rows=16):
    """Contiguous, `align`-row aligned bands with (nearly) equal summed cost.  `row_cost`: per-row cost estimate (len H).
    Every band keeps at least `min_rows` rows (the halo a neighbour reads from it must lie inside one band).
    Deterministic: every rank computes the same partition from the same costs."""
    row_cost = np.asarray(row_cost, dtype=np.float64)
    height = len(row_cost)
    units = math.ceil(height / align)
    mu = max(1, math.ceil(min_rows / align))
    if units < world_size * mu:
        return row_bands(height, world_size, align)
    unit_cost = np.add.reduceat(row_cost, np.arange(0, height, align))
    csum = np.concatenate([[0.0], np.cumsum(unit_cost)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        u = int(np.searchsorted(csum, target))
        # the cut closest to the target, leaving room for the remaining ranks
        if u > 0 and abs(csum[u - 1] - target) <= abs(csum[min(u, units)] - target):
            u -= 1
        u = max(cuts[-1] + mu, min(u, units - (world_size - r) * mu))
        cuts.append(u)
    cuts.append(units)
    return [(min(height, cuts[r] * align), min(height, cuts[r + 1] * align)) for r in range(world_size)]


def row_cost_from_features(features, background_weight=0.05):
    """Cost model for the row partition: a pixel whose camera ray meets the medium (total transmittance < 1 in the K0
    feature buffer) costs 1, a background pixel `background_weight` (it still runs K0/K1's traversal and the pass-through)."""
    tr = features["transmittance"]
    active = (tr != 1.0).sum(axis=1).astype(np.float64)
    return active + background_weight * tr.shape[1]


class _DevPtr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_view(ptr, nbytes, device):
    """Zero-copy uint8 tensor over a raw CUDA allocation owned by the pass."""
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


def exchange_row_halo(planes, band, halo, rank, world_size, bands=None, group=None, wait=True):
    """Exchange `halo` rows above/below `band` between neighbouring ranks.

    planes: list of 2-D (rows, row_bytes) uint8 tensors covering the FULL frame height (each rank holds valid data in
    its own band); after the call rows [r0-halo, r0) and [r1, r1+halo) hold the neighbours' data.
    wait=False returns the pending works instead of making the current stream wait for them (the caller overlaps the
    transfer with stages that do not read the halo).
    """
    r0, r1 = band
    ops = []
    H = planes[0].shape[0]
    for t in planes:
        if rank > 0:
            lo = max(0, r0 - halo)
            ops.append(dist.P2POp(dist.isend, t[r0:min(r1, r0 + halo)], rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, t[lo:r0], rank - 1, group))
        if rank < world_size - 1:
            hi = min(H, r1 + halo)
            ops.append(dist.P2POp(dist.isend, t[max(r0, r1 - halo):r1], rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, t[r1:hi], rank + 1, group))
    if not ops:
        return []
    # a neighbour's band can be shorter than the halo: both sides must agree on sizes, so clamp consistently
    works = dist.batch_isend_irecv(ops)
    if wait:
        for w in works:
            w.wait()
        return []
    return works


def gather_row_bands(planes, bands, rank, group=None):
    """Temporal-history fallback: make every rank's band of `planes` visible on all ranks.  One broadcast per
    band and plane (bands differ in size when cost-balanced).  Stream-ordered like the halo exchange, no host synchronisation."""
    works = []
    for t in planes:
        for r, (a, b) in enumerate(bands):
            if b > a:
                works.append(dist.broadcast(t[a:b], src=r, group=group, async_op=True))
    return works


def reprojection_row_bound(cam_prev, cam_cur, corners_world, height):
    """Upper bound, in image rows, on how far a world point inside the volume's bounding box can move between the previous and
    the current camera (K2 reads the history at the reprojected pixel, VR/TemporalReuse.cs.slang:168-209).  Both projections of
    the 8 box corners; a projective map moves interior points of the box at most ~ as far as its most-moved corner, a 25 % +
    2-row margin covers the perspective non-linearity.  Returns inf when a corner lies behind either camera."""
    def rows(cam):
        V = np.array(list(cam.viewMat), dtype=np.float64).reshape(4, 4)
        P = np.array(list(cam.projMat), dtype=np.float64).reshape(4, 4)
        c = np.concatenate([corners_world, np.ones((len(corners_world), 1))], axis=1) @ V @ P
        if (c[:, 3] <= 1e-6).any():
            return None
        return (-0.5 * c[:, 1] / c[:, 3] + 0.5) * height
    a, b = rows(cam_prev), rows(cam_cur)
    if a is None or b is None:
        return float("inf")
    return float(np.abs(a - b).max()) * 1.25 + 2.0


class ShardedPass:
    """Drives one ``VolumetricReSTIR`` pass per rank over its row band with halo exchanges between the stages."""

    def __init__(self, pass_, width, height, rank, world_size, device, temporal_halo=16):
        self.p = pass_
        self.W, self.H = width, height
        self.rank, self.world = rank, world_size
        self.device = device
        self.bands = row_bands(height, world_size)
        self.band = self.bands[rank]
        self.temporal_halo = temporal_halo
        min_band = min(b[1] - b[0] for b in self.bands)
        self.max_halo = min_band
        self._prev_cam = None
        self.history_gathers = 0      # frames whose history went through the all-gather fallback (motion beyond the halo)

    def balance(self, background_weight=0.05, refine=0, out_color_ptr=None):
        """Re-partition the rows by estimated cost (the scaling limiter of row sharding is load imbalance,
        sky rows are cheap).  Call after ``pass.setScene(scene, W, H)`` (full frame) and before the first frame.

        1. Every rank runs K0 over the whole frame once and derives the same per-row cost from the feature buffer
           (1 per pixel whose ray meets the medium, `background_weight` per background pixel) -> equal-cost bands.
        2. `refine` times (off by default: the fixed launch latency of short bands makes it over-correct): two frames are rendered on the current bands, the ranks all-gather their device time per frame,
           the model cost of every band is rescaled by its measured time and the rows are cut again (needs
           `out_color_ptr`, a full-frame float4 device buffer).
        The temporal history restarts afterwards.  Returns the band of this rank."""
        FEAT_DTYPE = np.dtype([("noReflectiveSurface", np.int32), ("transmittance", np.float32)])
        if self.world == 1:
            return self.band
        min_rows = max(16, self.temporal_halo, int(math.ceil(self.p.params.mSampleRadius)))
        self.p.setRowBand(0, self.H)
        self.p.execute_stage(0)
        feat = self.p.get_buffer(capi.BUF_FEATURES).view(FEAT_DTYPE).reshape(self.H, self.W)
        cost = row_cost_from_features(feat, background_weight)

        def apply(bands):
            self.bands = bands
            self.band = bands[self.rank]
            self.max_halo = min(b[1] - b[0] for b in bands)
            self.p.setRowBand(*self.band)
            self.p.updateDict({})          # a fresh option-change epoch: frame counter and history restart

        apply(balanced_row_bands(cost, self.world, min_rows=min_rows))
        for _ in range(refine if out_color_ptr else 0):
            for _f in range(2):
                self.execute(out_color_ptr)
            # compute time of this rank's band; the spatial stage's own event pair would include the wait for the neighbours' halo
            tm = self.p.timings()
            busy = tm["features_ms"] + tm["initial_ms"] + tm["temporal_ms"] + tm["final_ms"]
            try:
                mt = self.p.march_timings()
                busy += mt["spatial_cam_ms"] + mt["spatial_light_ms"]
            except capi.VRestirError:
                busy += tm["spatial_ms"]
            ms = torch.tensor([busy], dtype=torch.float64, device=self.device)
            every = [torch.zeros_like(ms) for _ in range(self.world)]
            dist.all_gather(every, ms)
            for r, (a, b) in enumerate(self.bands):
                s = cost[a:b].sum()
                if s > 0:
                    cost[a:b] *= float(every[r].item()) / s
            apply(balanced_row_bands(cost, self.world, min_rows=min_rows))
        return self.band

    def _planes(self, buffer):
        base, stride, planes = self.p.device_buffer(buffer)
        n = self.W * self.H
        out = []
        if buffer in (capi.BUF_RESERVOIR_0, capi.BUF_RESERVOIR_1, capi.BUF_RESERVOIR_TEMPORAL):
            for k in range(planes):
                out.append(device_view(base + k * stride, n * 16, self.device).view(self.H, self.W * 16))
        elif buffer in (capi.BUF_FEATURES, capi.BUF_FEATURES_TEMPORAL):
            out.append(device_view(base, n * 8, self.device).view(self.H, self.W * 8))
        elif buffer in (capi.BUF_PPARTIAL_0, capi.BUF_PPARTIAL_1, capi.BUF_PPARTIAL_TEMPORAL):
            out.append(device_view(base, n * 4, self.device).view(self.H, self.W * 4))
        else:
            B = self.p.params.mMaxBounces
            if B > 1 and base:
                out.append(device_view(base, n * (B - 1) * 12, self.device).view(self.H, self.W * (B - 1) * 12))
        return out

    def _with_ppartial(self, bufs):
        """Reservoir buffers travel with their p_partial plane when vertex reuse is on (VR/HostDeviceSharedDefinitions.h:29-31)."""
        prm = self.p.params
        if not (prm.mVertexReuse and prm.mMaxBounces > 1):
            return bufs
        twin = {capi.BUF_RESERVOIR_0: capi.BUF_PPARTIAL_0, capi.BUF_RESERVOIR_1: capi.BUF_PPARTIAL_1, capi.BUF_RESERVOIR_TEMPORAL: capi.BUF_PPARTIAL_TEMPORAL}
        return bufs + [twin[b] for b in bufs if b in twin]

    def _exchange(self, buffers, halo, wait=True):
        if self.world == 1:
            return []
        if int(halo) > self.max_halo:
            # a tap would land in rows that no neighbour ever sent: the sharded frame would silently differ from the single-GPU one
            raise capi.VRestirError(capi.ERR_INVALID_ARGUMENT, f"halo of {int(halo)} rows exceeds the smallest row band ({self.max_halo} rows): "
                                    "use fewer ranks, a smaller mSampleRadius or larger bands")
        planes = []
        for b in self._with_ppartial(list(buffers)):
            planes += self._planes(b)
        # NCCL send/recv are ordered after the pass's kernels through torch's current stream (the pass launches on the same,
        # default, stream) and work.wait() only makes that stream wait: no host synchronisation
        return exchange_row_halo(planes, self.band, halo, self.rank, self.world, wait=wait)

    def execute(self, out_color_ptr, out_mvec_ptr=None, stream=None):
        p = self.p
        prm = p.params
        B = prm.mMaxBounces
        radius = int(math.ceil(prm.mSampleRadius))
        if self.world == 1:   # nothing to exchange: the whole-frame call (K0 overlapped with K1 on an auxiliary stream)
            p.execute(out_color_ptr, out_mvec_ptr, stream)
            return
        p.execute_stage(0, 0, out_color_ptr, out_mvec_ptr, stream)
        p.execute_stage(1, 0, out_color_ptr, out_mvec_ptr, stream)
        p.execute_stage(7, 0, out_color_ptr, out_mvec_ptr, stream)   # "mPipelineFrames": K0/K1 of the next frame start now (else a no-op)
        for w in getattr(self, "_pending_history", []):   # the history halo of the previous frame travelled during K0/K1
            w.wait()
        self._pending_history = []
        if prm.mEnableTemporalReuse and not prm.mUseReference and self._history_needs_gather():
            # the camera of THIS frame (known only now) reprojects further than the halo that travelled, or the volume carries a
            # velocity field: every band of the history is made visible everywhere before K2 reads it
            B_ = prm.mMaxBounces
            bufs = [capi.BUF_RESERVOIR_TEMPORAL, capi.BUF_FEATURES_TEMPORAL] + ([capi.BUF_EXTRA_TEMPORAL] if B_ > 1 else [])
            planes = []
            for b in self._with_ppartial(bufs):
                planes += self._planes(b)
            for w in gather_row_bands(planes, self.bands, self.rank):
                w.wait()
            self.history_gathers += 1
        p.execute_stage(2, 0, out_color_ptr, out_mvec_ptr, stream)
        if prm.mEnableSpatialReuse and not prm.mUseReference:
            for r in range(prm.mSpatialReuseRounds):
                inb = p.spatial_input_buffer(r)
                bufs = [inb, capi.BUF_FEATURES] + ([capi.BUF_EXTRA_0 if inb == capi.BUF_RESERVOIR_0 else capi.BUF_EXTRA_1] if B > 1 else [])
                self._exchange(bufs[:1] + bufs[2:], radius)
                p.execute_stage(3, r, out_color_ptr, out_mvec_ptr, stream)
        p.execute_stage(4, 0, out_color_ptr, out_mvec_ptr, stream)
        p.execute_stage(5, 0, out_color_ptr, out_mvec_ptr, stream)
        p.execute_stage(6, 0, out_color_ptr, out_mvec_ptr, stream)
        if prm.mEnableTemporalReuse and not prm.mUseReference:
            # history for the next frame's K2.  Reprojected taps normally land within `temporal_halo` rows of the band (halo
            # exchange with the two neighbours); when the camera moved further than that between the last two frames (the best
            # predictor of the next step), or the volume carries a velocity field, every band is made visible everywhere.
            bufs = [capi.BUF_RESERVOIR_TEMPORAL, capi.BUF_FEATURES_TEMPORAL] + ([capi.BUF_EXTRA_TEMPORAL] if B > 1 else [])
            self._pending_history = self._exchange(bufs, min(self.temporal_halo, self.max_halo), wait=False)
            self._prev_cam = self.p._scene.camera.data(self.W, self.H)      # the camera this history was rendered with

    def _history_needs_gather(self):
        """Called at the start of a frame, when its camera is known: can K2's reprojected taps leave band + halo?"""
        sc = self.p._scene
        if self._prev_cam is None or self.p.frame_count() == 0:
            return False                                   # no history yet (first frame of an option epoch)
        vd = sc.volume.grid.contents.volume
        if vd.hasVelocity and vd.hasAnimation:
            return True                                    # the velocity field moves the reprojection point arbitrarily
        lo, hi = sc.volume_bounds_world()
        corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])], dtype=np.float64)
        return reprojection_row_bound(self._prev_cam, sc.camera.data(self.W, self.H), corners, self.H) > min(self.temporal_halo, self.max_halo)
