"""Multi-GPU product entry (python -m pyp_b200.run): the maps, accumulators and parameter files written by a
2-GPU launch (contiguous particle shards, NCCL sum of the half-volume accumulators on rank 0) equal those of a
single-GPU launch over the union.  Needs two visible GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import small_case
from oracle import tier_a
from pyp_b200 import synth
from pyp_b200.formats import cistem, dump, mrc, statistics

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from pyp_b200 import _lib

    return int(_lib.lib().cspb_device_count())


def _launch(gpus, d, extra=()):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "pyp_b200.run", "--gpus", str(gpus), "--refine", "refine.in", "--reconstruct", "reconstruct.in",
           "--merge", "merge.in", "--keep-dumps", "--out-log", "run.log", *extra]
    r = subprocess.run(cmd, cwd=d, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return open(os.path.join(d, "run.log")).read()


def _dataset(d, n=64, px=1.35, P=240):
    ph, vol, rows, stack = small_case(n=n, n_part=P, snr=0.5)
    start = synth.perturb_rows(rows, 2.0, 1.0)
    os.makedirs(d, exist_ok=True)
    mrc.write(f"{d}/ds_stack.mrc", stack, px)
    mrc.write(f"{d}/ds_r01.mrc", vol, px)
    cistem.write_parameters(f"{d}/ds_r01.cistem", start)
    open(f"{d}/statistics_r01.txt", "w").close()
    a = tier_a.refine3d_answers("ds_stack.mrc", "ds_r01.cistem", "ds_r01.mrc", "statistics_r01.txt", "ds_r01", 1, P, px, 100.0, 0.38 * n * px, 60.0, 4 * px)
    open(f"{d}/refine.in", "w").write(tier_a.heredoc(a))
    a = tier_a.reconstruct3d_answers("ds_stack.mrc", "ds_r01_used.cistem", "ds_r01.mrc", "ds_r01", 1, P, px, 100.0, px * n / 2, 2 * px,
                                     "ds_r01_map1_n1.mrc", "ds_r01_map2_n1.mrc")
    open(f"{d}/reconstruct.in", "w").write(tier_a.heredoc(a))
    a = tier_a.merge3d_answers("ds_r01_02", 100.0, px * n / 2, "ds_r01_map1_n.mrc", "ds_r01_map2_n.mrc", 1)
    open(f"{d}/merge.in", "w").write(tier_a.heredoc(a))
    return rows, start, vol


def test_single_gpu_run_writes_pyps_file_set(tmp_path):
    d = str(tmp_path / "one")
    rows, start, vol = _dataset(d)
    log = _launch(1, d)
    assert "Refine3D: Normal termination" in log and "Merge3D: Normal termination" in log
    out = cistem.read_parameters(f"{d}/ds_r01_0000001_0000240.cistem")
    assert list(out["position_in_stack"]) == list(range(1, 241)) and os.path.exists(f"{d}/ds_r01_0000001_0000240_changes.cistem")
    from common import angular_distance

    assert angular_distance(out, rows).mean() < angular_distance(start, rows).mean()
    table = statistics.parse_merge3d_log(log)  # frealign.py:2558-2567
    assert table.shape[1] == 7 and table[1:5, 3].min() > 0.5
    for name in ("ds_r01_02.mrc", "ds_r01_02_half1.mrc", "ds_r01_02_half2.mrc", "ds_r01_02_statistics.txt"):
        assert os.path.exists(f"{d}/{name}")


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_run_equals_single_gpu_run(tmp_path):
    from oracle import oracle as O

    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    _dataset(one)
    _dataset(two)
    _launch(1, one, ("--cutoff", "0.8"))
    log2 = _launch(2, two, ("--cutoff", "0.8"))
    assert "on 2 GPU(s)" in log2
    # refine3d: per-particle work with one broadcast whitening curve -> the same rows
    r1 = cistem.read_parameters(f"{one}/ds_r01_0000001_0000240.cistem")
    r2 = cistem.read_parameters(f"{two}/ds_r01_0000001_0000240.cistem")
    for k in ("psi", "theta", "phi", "x_shift", "y_shift", "score", "sigma", "logp"):
        assert np.allclose(r1[k], r2[k], rtol=1e-6, atol=1e-6), k
    # reconstruct3d: rank-0 accumulators after the NCCL reduce == single-GPU accumulators over the union
    for h in (1, 2):
        m1, a1 = dump.read(f"{one}/ds_r01_map{h}_n1.mrc")
        m2, a2 = dump.read(f"{two}/ds_r01_map{h}_n1.mrc")
        assert m1["n_inserted"] == m2["n_inserted"] == 240 - int(239 * (1 - 0.8))  # the 47 lowest scores dropped (scores.py:486-497)
        assert np.abs(a1 - a2).max() <= 1e-5 * np.abs(a1).max()
    # merge3d: maps FSC >= 0.999 at every shell, statistics equal
    for name in ("ds_r01_02.mrc", "ds_r01_02_half1.mrc", "ds_r01_02_half2.mrc"):
        v1, v2 = np.asarray(mrc.read(f"{one}/{name}")[1]), np.asarray(mrc.read(f"{two}/{name}")[1])
        assert O.fsc(v1, v2)[1:].min() >= 0.999
    s1, s2 = statistics.read_statistics(f"{one}/ds_r01_02_statistics.txt"), statistics.read_statistics(f"{two}/ds_r01_02_statistics.txt")
    assert np.abs(s1[:, 3] - s2[:, 3]).max() < 1e-3
