import sys, numpy as np
sys.path.insert(0, '/root/repo')
# ---- worst row deviation of the fast projection ----
import lt_mapper_b200 as ltr
rng = np.random.default_rng(42)
def _random_pose(rng, big=False):
    T = np.eye(4)
    qq = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    if np.linalg.det(qq) < 0:
        qq[:, 0] = -qq[:, 0]
    T[:3, :3] = qq
    T[:3, 3] = rng.uniform(-500, 500, 3) if big else rng.uniform(-60, 60, 3)
    return T
with ltr.Context() as ctx:
    for trial in range(2):
        pose = _random_pose(rng, big=trial % 2 == 1)
        inv = np.linalg.inv(pose)
        n = 2_000_000
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        d[: n // 8, :2] *= 1e-3
        d[n // 8: n // 4, 1] = rng.uniform(-1e-4, 1e-4, n // 8); d[n // 8: n // 4, 0] = -np.abs(d[n // 8: n // 4, 0])
        r = np.exp(rng.uniform(np.log(0.5), np.log(150.0), n))[:, None]
        local = d * r
        world = (local @ pose[:3, :3].T + pose[:3, 3]).astype(np.float32)
        out, mg = ctx.debug_fast_project(world, inv, 2.5)
        t = local[:, 2] / np.hypot(local[:, 0], local[:, 1])
        inside = (np.abs(t) < 0.4995) & np.isfinite(out).all(axis=1)
        drow = np.abs(out[:, 1].astype(np.float64) - out[:, 5])
        drow[~inside] = 0
        i = int(np.argmax(drow))
        print("trial", trial, "worst drow", drow[i], "margin", mg[2], "t", t[i], "local", local[i], "fast vrow", out[i, 1], "ref vrow", out[i, 5], "range", out[i, 2], out[i, 6])
        for i in np.argsort(-drow)[:5]:
            print("     i", i, "drow", drow[i], "t", t[i], "local", local[i], "world", world[i], "fast", out[i, :4], "ref", out[i, 4:7])
        print("   fraction with drow > margin:", float((drow > mg[2]).mean()), "margins", ctx.debug_margins(2.5))
