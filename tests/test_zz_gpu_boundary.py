"""Rounding-sensitive inputs on the device (added after round 1's GPU budget was spent, so it sorts last: a failure here must not hide
the rest of the suite)."""
import numpy as np
import pytest

import oracle
import lt_mapper_b200 as ltr

pytestmark = pytest.mark.gpu


def test_points_on_pixel_boundaries():
    """Points whose pre-round pixel coordinate is n + 0.5 and their float neighbours (the inputs on which the fast pixel evaluation must
    hand over to the exact arithmetic): device pixel index == oracle (pinned against the compiled reference on the same construction in
    tests/test_ref_pin.py), and a scan image built from them, fast path on, == the oracle's image."""
    rng = np.random.default_rng(5)
    for (vfov, hfov, alpha), count in (((50.0, 360.0, 0.4), 20000), ((50.0, 360.0, 2.5), 20000), ((33.2, 180.0, 1.0), 20000)):
        rows, cols = oracle.reset_rimg_size(alpha, vfov, hfov)
        r = rng.uniform(0.5, 90.0, count)
        colb = rng.random(count) < 0.5
        az = np.where(colb, np.deg2rad((rng.integers(0, cols, count) + 0.5) / cols * hfov - hfov / 2), np.deg2rad(rng.uniform(-hfov / 2, hfov / 2, count)))
        el = np.where(colb, np.deg2rad(rng.uniform(-vfov / 2, vfov / 2, count)), np.deg2rad(vfov / 2 - (rng.integers(0, rows, count) + 0.5) / rows * vfov))
        pts = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], 1).astype(np.float32)
        for _ in range(3):               # walk 0-3 ulps in random coordinates
            j = rng.integers(0, 3, count); step = rng.integers(-1, 2, count)
            cur = pts[np.arange(count), j]
            pts[np.arange(count), j] = np.where(step == 0, cur, np.nextafter(cur, np.where(step > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32)))
        er, ec, erng = oracle.pixel_index(pts, rows, cols, vfov, hfov)
        xyzi = np.concatenate([pts, np.zeros((count, 1), np.float32)], 1)
        exp_img = oracle.scan2rimg(xyzi, rows, cols, vfov, hfov)
        for fast in (True, False):
            with ltr.Context(vfov=vfov, hfov=hfov, fast_path=fast) as ctx:
                gr, gc, grng, _, _ = ctx.debug_pixel_index(pts, rows, cols)
                assert np.array_equal(gr, er) and np.array_equal(gc, ec) and np.array_equal(grng.view(np.uint32), erng.view(np.uint32))
                img = ctx.debug_scan_rimg(ctx.scanset_upload(xyzi, [0, count]), 0, alpha)
                assert np.array_equal(img.view(np.uint32), exp_img.view(np.uint32)), (vfov, alpha, fast)
