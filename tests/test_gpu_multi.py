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


def test_standalone_driver_on_two_gpus_writes_the_single_process_tree(tmp_path, small_pair):
    """apps/ltremovert_b200 started once per GPU (RANK / WORLD_SIZE / LOCAL_RANK, NCCL id through a file -- no Python, no torch in the
    processes): every rank loads only its keyframe block from disk, and the files written are those of the single-process run."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import filecmp
    from test_gpu_driver import BIN, _write_session
    c, q = small_pair
    _write_session(tmp_path / "central", c)
    _write_session(tmp_path / "query", q)

    def cfg(out):
        p = tmp_path / f"params_{out}.yaml"
        p.write_text("removert:\n  saveMapPCD: true\n" + f'  save_pcd_directory: "{tmp_path / out}"\n'
                     + f'  central_sess_scan_dir: "{tmp_path}/central/Scans/"\n  central_sess_pose_path: "{tmp_path}/central/poses.txt"\n'
                     + f'  query_sess_scan_dir: "{tmp_path}/query/Scans/"\n  query_sess_pose_path: "{tmp_path}/query/poses.txt"\n'
                     + "  sequence_vfov: 50\n  sequence_hfov: 360\n  keyframe_gap: 1\n  start_idx: 0\n  end_idx: 5\n  remove_resolution_list: [2.5, 2.0]\n"
                     + "  downsample_voxel_size: 0.05\n  num_nn_points_within: 2\n  dist_nn_points_within: 0.01\n")
        return str(p)
    r = subprocess.run([BIN, "--config", cfg("one"), "--selfremovert"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for mode, extra in (("two_split", []), ("two_blocks", ["--no-split"])):
        procs = []
        for rank in range(2):
            env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_PORT="29601", LTR_NCCL_ID_FILE=str(tmp_path / f"id_{mode}"),
                       LTR_DIST_VOXEL_MIN="1000")
            procs.append(subprocess.Popen([BIN, "--config", cfg(mode), "--selfremovert"] + extra, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)

        def tree(d):
            return sorted(os.path.relpath(os.path.join(p, f), d) for p, _, fs in os.walk(d) for f in fs)
        files = tree(tmp_path / "one")
        assert files == tree(tmp_path / mode) and len(files) >= 14 + 5 * 3, (mode, sorted(set(files) ^ set(tree(tmp_path / mode))))
        for f in files:
            assert filecmp.cmp(tmp_path / "one" / f, tmp_path / mode / f, shallow=False), (mode, f)
