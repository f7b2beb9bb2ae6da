"""Driver-visible multi-GPU parity ("identical ND/PD index sets for every N"): when the box has >= 2 GPUs, 2 (and 4)
ranks over NCCL run Steps 0-3 on a small pair, session-split and not, and every map, every saved cloud and every per-keyframe cloud
must equal the single-process run byte for byte.  The workers go through the same C-ABI as everything else."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(n, out, split, kf=8, cascade=0):
    worker = os.path.join(ROOT, "tests", "_multi_rank_worker.py")
    args = ["--out", out, "--kf", str(kf), "--split", str(split), "--cascade", str(cascade)]
    if n == 1:
        cmd = [sys.executable, worker] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + 7 * n + split), worker] + args
    env = dict(os.environ, LTR_DIST_VOXEL_MIN="1000")   # the small pair must exercise the distributed voxelisation of appended clouds too
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    merged = {}
    for f in sorted(glob.glob(os.path.join(out, "rank*.json"))):
        d = json.load(open(f))
        d.pop("__log__")
        for k, v in d.items():
            assert merged.setdefault(k, v) == v, f"replicas disagree on {k}"
    return merged


@pytest.mark.parametrize("n,split", [(2, 1), (2, 0), (4, 1)])
def test_multi_rank_equals_single_rank(tmp_path, n, split):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    one = _run(1, str(tmp_path / "one"), 0)
    many = _run(n, str(tmp_path / f"n{n}s{split}"), split)
    assert len(one) > 100
    missing = sorted(set(one) - set(many))
    assert not missing, missing[:10]
    bad = [k for k in one if many[k] != one[k]]
    assert not bad, bad[:10]


def test_multi_rank_cascade_equals_single_rank(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    one = _run(1, str(tmp_path / "one"), 0, kf=6, cascade=1)
    two = _run(2, str(tmp_path / "two"), 1, kf=6, cascade=1)
    bad = [k for k in one if two.get(k) != one[k]]
    assert not bad, bad[:10]
