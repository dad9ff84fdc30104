"""Host-side mirror of the reference's pass interface: option surface, defaults, dictionary semantics, row bands."""
import os
import re

import numpy as np
import pytest

from volumetricrestirrelease_b200 import VolumetricReSTIRParams, capi
from volumetricrestirrelease_b200.multi_gpu import row_bands

REF_HEADER = "/root/reference/Source/RenderPasses/VolumetricReSTIR/VolumetricReSTIR.h"


def test_params_defaults_and_unknown_field():
    p = VolumetricReSTIRParams()
    assert p.mInitialM == 4 and p.mSpatialSampleCount == 4 and p.mSampleRadius == 10.0 and p.mMaxBounces == 1
    assert p.mFinalVisibilityTrackingMethod == capi.kAnalyticTracking and p.mSpatialMISMethod == capi.kMISTalbot
    with pytest.raises(AttributeError):
        VolumetricReSTIRParams(mNotAField=1)
    c = VolumetricReSTIRParams(mInitialM=7, mFinalTStepScale=0.5).to_c()
    q = VolumetricReSTIRParams.from_c(c)
    assert q.mInitialM == 7 and abs(q.mFinalTStepScale - 0.5) < 1e-7


@pytest.mark.skipif(not os.path.exists(REF_HEADER), reason="reference checkout absent")
def test_option_surface_matches_reference_header():
    """Every field of VolumetricReSTIRParams in the reference header exists here with the same default."""
    text = open(REF_HEADER).read()
    body = text[text.index("struct VolumetricReSTIRParams"):text.index("} mParams;")]
    consts = {"kRayMarching": 2, "kAnalyticTracking": 1, "kReprojectionLinear": 0, "kMISTalbot": 1, "kR2": 1, "true": 1, "false": 0}
    found = {}
    for m in re.finditer(r"(?:int|bool|float|uint32_t)\s+(m\w+)\s*=\s*([^;]+);", body):
        v = m.group(2).strip().rstrip("f")
        found[m.group(1)] = float(consts.get(v, v if v[-1] != "." else v + "0"))
    mine = {n: float(d) for n, _, d in capi.PARAM_FIELDS}
    assert set(found) == set(mine), set(found) ^ set(mine)
    for k, v in found.items():
        assert abs(mine[k] - v) < 1e-6, k


def test_row_bands_cover_frame_and_are_tile_aligned():
    for h in (1080, 2160, 96, 100):
        for n in (1, 2, 4, 8):
            b = row_bands(h, n)
            assert b[0][0] == 0 and b[-1][1] == h
            assert all(b[i][1] == b[i + 1][0] for i in range(n - 1))
            assert all(r0 % 8 == 0 for r0, _ in b)
            sizes = [r1 - r0 for r0, r1 in b]
            assert max(sizes) - min(sizes) <= 8 + (h % 8)


def test_balanced_row_bands_partition_and_balance():
    """Cost-balanced bands: contiguous, aligned, cover the frame, respect the minimum band height, and are better balanced
    than the even split for a frame whose cost sits in the middle rows (cloud in front of sky)."""
    import numpy as np
    from volumetricrestirrelease_b200.multi_gpu import balanced_row_bands, row_bands
    H = 1080
    y = np.arange(H)
    cost = 0.05 * 1920 + 1920 * np.exp(-((y - 520) / 180.0) ** 2)
    for n in (2, 3, 4, 8):
        bands = balanced_row_bands(cost, n)
        assert bands[0][0] == 0 and bands[-1][1] == H
        assert all(b[1] == bands[i + 1][0] for i, b in enumerate(bands[:-1]))
        assert all(b[0] % 8 == 0 for b in bands) and all(b[1] - b[0] >= 16 for b in bands)
        even = row_bands(H, n)
        worst = lambda bs: max(cost[a:b].sum() for a, b in bs)
        assert worst(bands) <= worst(even) + 1e-9
        if n >= 4:
            assert worst(bands) < 0.75 * worst(even)
    # degenerate: fewer aligned units than ranks x minimum -> even split
    assert balanced_row_bands(np.ones(40), 4) == row_bands(40, 4)
