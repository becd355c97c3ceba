"""End-to-end drop-in test: drive bin/{refine3d,reconstruct3d,local_merge3d,merge3d} the way pyp
does — `sh -c "<prog> << eot ... eot"` with the answer lists of frealign.py:3918-3994,
1780-1824, 1878-1888, 2075-2093 — on a small synthetic data set, then check the files pyp
looks for and the log table it parses."""
import os
import subprocess

import numpy as np
import pytest

from common import angular_distance, small_case
from pyp_b200 import synth
from pyp_b200.formats import cistem, mrc, statistics

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def sh(prog, answers, cwd, log):
    cmd = f"{BIN}/{prog} << eot >> {log} 2>&1\n" + "\n".join(str(a) for a in answers) + "\neot\n"
    return subprocess.run(cmd, shell=True, cwd=cwd, timeout=600).returncode


def test_refine_reconstruct_merge_chain(tmp_path):
    n, px, n_part = 64, 1.35, 60
    ph, vol, rows, stack = small_case(n=n, n_part=n_part, snr=0.5)
    start = synth.perturb_rows(rows, 2.0, 1.0)
    d = str(tmp_path)
    mrc.write(f"{d}/ds_stack.mrc", stack, px)
    mrc.write(f"{d}/ds_r01.mrc", vol, px)
    cistem.write_parameters(f"{d}/ds_r01.cistem", start)
    open(f"{d}/statistics_r01.txt", "w").close()
    outs = []
    for first, last in [(1, 30), (31, 60)]:  # two concurrent ranges like local_run.py:507-516
        ranger = "%07d_%07d" % (first, last)
        a = ["ds_stack.mrc", "ds_r01.cistem", "null", "ds_r01.mrc", "statistics_r01.txt", "no", "no", f"ds_r01_match.mrc_{ranger}",
             f"ds_r01_{ranger}.cistem", f"ds_r01_{ranger}_changes.cistem", "C1", first, last, 1, px, 100.0, 0, 0.38 * n * px, 60.0, 4 * px,
             "30.0", 8.0, 1.5 * 0.38 * n * px, 4 * px, 20.0, 20, 0, 0, 0, 0, 0, 0, 500, "50.0", 1, "no", "yes", "yes", "yes", "yes", "yes",
             "yes", "no", "no", "no", "yes", "no", "no", "no", "no"]
        assert sh("refine3d", a, d, "refine.log") == 0
        outs.append(f"{d}/ds_r01_{ranger}.cistem")
        assert os.path.exists(outs[-1]) and os.path.exists(outs[-1].replace(".cistem", "_changes.cistem"))
    assert "Refine3D: Normal termination" in open(f"{d}/refine.log").read()
    refined = cistem.merge(outs)  # Parameters.merge semantics
    assert list(refined["position_in_stack"]) == list(range(1, n_part + 1))
    assert angular_distance(refined, rows).mean() < angular_distance(start, rows).mean()
    assert (refined["score"] > 0).all() and (refined["score"] <= 100).all()
    cistem.write_parameters(f"{d}/ds_r01_used.cistem", refined)

    os.makedirs(f"{d}/scratch", exist_ok=True)
    for k, (first, last) in enumerate([(1, 20), (21, 40), (41, 60)], start=1):
        a = ["ds_stack.mrc", "ds_r01_used.cistem", "null", "ds_r01.mrc", "ds_r01_map1.mrc", "ds_r01_map2.mrc", "output.mrc", f"ds_r01_n{first}.res",
             "C1", first, last, px, 100.0, 0, px * n / 2, 2 * px, 0, 2.0, "no", 0, -1, "no", 0, 1, 1, "yes", "no", "no", "no", "no", "yes", "no",
             "no", "no", "no", "yes", f"scratch/ds_r01_map1_n{k}.mrc", f"scratch/ds_r01_map2_n{k}.mrc", 1]
        assert sh("reconstruct3d", a, d, "recon.log") == 0
    assert "caught" not in open(f"{d}/recon.log").read()  # particle_cspt.py:812-818

    # local_merge3d over the first two, as local_merge_reconstruction renames them (frealign.py:1870-1888)
    os.rename(f"{d}/scratch/ds_r01_map1_n1.mrc", f"{d}/temp_map1_n1.mrc")
    os.rename(f"{d}/scratch/ds_r01_map2_n1.mrc", f"{d}/temp_map2_n1.mrc")
    os.rename(f"{d}/scratch/ds_r01_map1_n2.mrc", f"{d}/temp_map1_n2.mrc")
    os.rename(f"{d}/scratch/ds_r01_map2_n2.mrc", f"{d}/temp_map2_n2.mrc")
    assert sh("local_merge3d", ["dumpfile_map1.mrc", "dumpfile_map2.mrc", "temp_map1_n.mrc", "temp_map2_n.mrc", 2], d, "local_merge3d.log") == 0
    assert os.path.exists(f"{d}/dumpfile_map1.mrc")  # frealign.py:1892
    os.rename(f"{d}/dumpfile_map1.mrc", f"{d}/scratch/ds_r01_map1_n1.mrc")
    os.rename(f"{d}/dumpfile_map2.mrc", f"{d}/scratch/ds_r01_map2_n1.mrc")
    os.rename(f"{d}/scratch/ds_r01_map1_n3.mrc", f"{d}/scratch/ds_r01_map1_n2.mrc")
    os.rename(f"{d}/scratch/ds_r01_map2_n3.mrc", f"{d}/scratch/ds_r01_map2_n2.mrc")
    a = ["ds_r01_02_half1.mrc", "ds_r01_02_half2.mrc", "ds_r01_02.mrc", "ds_r01_02_statistics.txt", 100.0, 0, px * n / 2,
         "scratch/ds_r01_map1_n.mrc", "scratch/ds_r01_map2_n.mrc", 2]
    assert sh("merge3d", a, d, "merge.log") == 0
    table = statistics.parse_merge3d_log(open(f"{d}/merge.log").read())  # frealign.py:2558-2567
    assert table.shape == (n // 2, 7) and table[0, 1] == pytest.approx(n * px, abs=0.01)
    assert table[1:5, 3].min() > 0.5
    _, m = mrc.read(f"{d}/ds_r01_02.mrc")
    assert m.shape == (n, n, n)
    from oracle import oracle as O

    f = O.fsc(np.asarray(m), vol)
    assert f[1:8].min() > 0.8
    st = statistics.read_statistics(f"{d}/ds_r01_02_statistics.txt")
    assert st.shape == (n // 2, 7)

    # error convention: bad input -> non-zero exit and the word "caught" in the log
    assert sh("reconstruct3d", ["missing.mrc"], d, "bad.log") != 0
    assert "caught" in open(f"{d}/bad.log").read()
