"""CPU: the accumulate / error-measure oracle against hand-computed known answers (the reference holds no fixtures for these passes)."""
import numpy as np

from oracle import post_oracle as po


def test_accumulate_known_answers():
    frames = [np.full((2, 2, 4), v, np.float32) for v in (1.0, 2.0, 6.0)]
    for mode in ("Single", "SingleCompensated", "Double"):
        acc = po.Accumulator(mode)
        outs = [acc.add(f) for f in frames]
        assert np.all(outs[0] == 1.0) and np.all(outs[1] == 1.5) and np.all(outs[2] == 3.0), mode
        assert acc.count == 3


def test_compensated_beats_single_on_long_sums():
    rng = np.random.default_rng(0)
    vals = (rng.random((4000, 1, 1, 4)) * 1e-3 + 1.0).astype(np.float32)
    exact = vals.astype(np.float64).mean(axis=0)
    errs = {}
    for mode in ("Single", "SingleCompensated", "Double"):
        acc = po.Accumulator(mode)
        for v in vals:
            out = acc.add(v)
        errs[mode] = np.abs(out - exact).max()
    assert errs["SingleCompensated"] <= errs["Single"] and errs["Double"] <= errs["Single"]
    assert errs["SingleCompensated"] < 2e-7 and errs["Double"] < 2e-7


def test_error_measure_known_answers():
    src = np.zeros((2, 2, 4), np.float32)
    ref = np.zeros((2, 2, 4), np.float32)
    src[0, 0, :3] = (1.0, 2.0, 3.0)            # one differing pixel
    src[1, 1, :3] = (0.5, 0.0, 0.0)            # a background pixel (world w == 0)
    wp = np.ones((2, 2, 4), np.float32); wp[1, 1, 3] = 0.0
    d, err, avg = po.error_measure(src, ref, wp, ignore_background=True, squared=True, average=False)
    assert np.array_equal(err, np.float32([1.0, 4.0, 9.0]) / np.float32(4)) and avg == np.float32((0.25 + 1.0 + 2.25) / 3)
    d, err, avg = po.error_measure(src, ref, wp, ignore_background=False, squared=False, average=False)
    assert np.allclose(err, [1.5 / 4, 2.0 / 4, 3.0 / 4])
    d, err, avg = po.error_measure(src, ref, None, ignore_background=True, squared=False, average=True)   # unbound world position: no background test
    assert np.allclose(d[0, 0], 2.0) and np.allclose(d[1, 1], 0.5 / 3) and np.allclose(err, (2.0 + 0.5 / 3) / 4)


def test_tonemap_color_transform_host_side():
    """ToneMapper::updateColorTransform / calculateWhiteBalanceTransformRGB_Rec709 through the ABI's host-only entry point
    against the numpy restatement; the D65 white point is preserved exactly at 6500 K (ColorUtils.h:190-193)."""
    import ctypes as C
    from oracle import post_oracle as po
    from volumetricrestirrelease_b200 import capi
    from volumetricrestirrelease_b200.post import ToneMapper
    for kw in ({}, {"whiteBalance": True, "whitePoint": 3200.0}, {"whiteBalance": True, "whitePoint": 9000.0, "exposureCompensation": 1.5},
               {"fNumber": 2.8, "shutter": 60.0, "filmSpeed": 400.0}, {"autoExposure": True, "exposureCompensation": -1.0}):
        tm = ToneMapper(kw)
        M = np.array(list(tm.params().colorTransform), dtype=np.float64).reshape(3, 3).T      # stored transposed for the row-vector product
        want = po.tonemap_color_transform(**kw)
        np.testing.assert_allclose(M, want, rtol=2e-6, atol=1e-7)
    tm = ToneMapper({"whiteBalance": True, "whitePoint": 6500.0})
    M = np.array(list(tm.params().colorTransform), dtype=np.float64).reshape(3, 3).T
    np.testing.assert_allclose(M @ np.ones(3), np.ones(3), rtol=1e-5)
    assert abs(ToneMapper({"exposureValue": 5.0}).exposureValue - 5.0) < 1e-5
