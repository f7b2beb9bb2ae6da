"""Stress test of the tile culling + fast path: maps that are NOT consistent with the scans (random blobs, shells around the
sensor, points at the keyframe origins, shuffled order = huge tiles, voxel order = small tiles), several resolutions and
thresholds.  Flags must equal the all-exact kernel (and the oracle) in every configuration."""
import numpy as np
import pytest

import oracle
import lt_mapper_b200 as ltr

pytestmark = pytest.mark.gpu


def _maps(rng, c):
    inv = oracle.inverse_poses(c.poses)
    origins = c.poses[:, :3, 3].astype(np.float32)
    merged = np.concatenate([oracle.transform(c.scan(k), c.poses[k]) for k in range(c.K)])
    lo, hi = merged[:, :3].min(0), merged[:, :3].max(0)
    out = {}
    # (a) the real map in voxel (Morton) order: small tiles, heavy culling
    out["voxel"] = oracle.voxel(merged, 0.1)
    # (b) the same points shuffled: every tile spans the whole scene
    out["shuffled"] = out["voxel"][rng.permutation(len(out["voxel"]))]
    # (c) random blobs in front of / behind the real surfaces + uniform clutter, voxel-ordered
    blobs = [merged[rng.integers(len(merged)), :3] + rng.normal(0, s, (4000, 3)) for s in (0.05, 0.3, 2.0) for _ in range(6)]
    clutter = rng.uniform(lo - 5, hi + 5, (60000, 3))
    pts = np.concatenate(blobs + [clutter]).astype(np.float32)
    out["blobs"] = oracle.voxel(np.concatenate([pts, np.zeros((len(pts), 1), np.float32)], 1), 0.05)
    # (d) thin shells around the sensor origins (ranges just inside / outside the scan ranges) and points at the origins themselves
    d = rng.normal(size=(50000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    shells = np.concatenate([origins[rng.integers(c.K)] + d[i::5] * r for i, r in enumerate((0.01, 2.4, 9.9, 30.0, 99.0))])
    near = np.concatenate([origins, origins + 1e-4, origins + rng.normal(0, 1e-3, origins.shape)]).astype(np.float32)
    pts = np.concatenate([shells, near]).astype(np.float32)
    out["shells"] = oracle.voxel(np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1), 0.02)
    # (e) the real surfaces pushed radially away from one keyframe origin (occluded there: the case culling is for),
    #     mixed with the same surfaces pulled towards it (every such point is a candidate)
    v = out["voxel"][:, :3]
    o = origins[rng.integers(c.K, size=len(v))]
    s = np.where(rng.random(len(v)) < 0.8, 1.0 + rng.uniform(0.02, 0.5, len(v)), 1.0 - rng.uniform(0.02, 0.5, len(v)))[:, None]
    pts = (o + (v - o) * s).astype(np.float32)
    out["behind"] = oracle.voxel(np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1), 0.1)
    return out, inv


@pytest.mark.parametrize("alpha,thres", [(2.5, 0.1), (1.0, 0.1), (3.0, 0.5), (0.5, 0.02)])
def test_culling_never_changes_flags(small_pair, alpha, thres):
    rng = np.random.default_rng(123)
    c = small_pair[0]
    maps, inv = _maps(rng, c)
    for name, m in maps.items():
        res = {}
        culled = 0.0
        for key, (fast, cull) in {"exact": (False, False), "fast": (True, False), "cull": (True, True)}.items():
            with ltr.Context(fast_path=fast, cull=cull) as ctx:
                mh = ctx.cloud_upload(m); ss = ctx.scanset_upload(c.xyzi, c.offsets); ps = ctx.poses_upload(c.poses, inv)
                out = []
                for mode in (ltr.MODE_HD, ltr.MODE_PD):
                    ctx.remove_pass(mh, ss, ps, mode, alpha, diff_thres=thres)
                    out.append(ctx.flags_download(mh))
                    if key == "cull":
                        st = ctx.last_pass_stats(); culled = max(culled, st[6] / max(st[0], 1))
                res[key] = out
        for i in range(2):
            assert np.array_equal(res["exact"][i], res["fast"][i]), (name, i)
            assert np.array_equal(res["exact"][i], res["cull"][i]), (name, i)
        exp = oracle.remove_pass(m, c.xyzi, c.offsets, inv, oracle.MODE_HD, alpha, thres)
        assert np.array_equal(res["cull"][0], exp), name
        print(f"alpha {alpha} thres {thres} map {name}: N={len(m)} flagged={int(exp.sum())} culled share up to {culled:.3f}")
