"""Host-side logic that needs no GPU: prompt parsing of the drop-in front-ends, range
partitioning, and the multi-process (gloo, world_size 2) gather / reduce plumbing."""
import os
import subprocess
import sys

import numpy as np
import pytest

from pyp_b200 import dist as pd
from pyp_b200._lib import ROW_DTYPE as ROW_DTYPE_
from pyp_b200.cli import local_merge3d, merge3d, prompts, reconstruct3d, refine3d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _refine3d_heredoc(first=1, last=100, sym="O", mode_local=True):
    """Answer list exactly as frealign.mrefine_version assembles it (frealign.py:3918-3994)."""
    mask = ["yes"] * 5
    lines = ["../T20S_stack.mrc", "T20S_r01.cistem", "null", "T20S_r01.mrc", "statistics_r01.txt", "no", "no",
             "T20S_r01_match.mrc_0000001_0000100", "T20S_r01_0000001_0000100.cistem", "T20S_r01_0000001_0000100_changes.cistem",
             sym, str(first), str(last), "1", "1.35", "700.0", "0", "80", "100.0", "8.0", "30.0", "8.0", "120.0", "8.0", "20.0", "20",
             "0", "0", "0", "0", "0", "0", "500", "50.0", "1", "no" if mode_local else "yes", "yes" if mode_local else "no"] + mask + \
            ["no", "no", "no", "yes", "no", "no", "no", "no"]
    return "\n".join(lines) + "\n"


def test_refine3d_prompt_order():
    p = refine3d.parse(prompts.Answers(_refine3d_heredoc(), "refine3d"))
    assert p["stack"] == "../T20S_stack.mrc" and p["out_changes"].endswith("_changes.cistem")
    assert (p["first"], p["last"], p["symmetry"]) == (1, 100, "O")
    assert p["pixel_size"] == 1.35 and p["outer_mask_radius"] == 80 and p["low_res_limit"] == 100.0
    assert p["high_res_limit"] == 8.0 and p["signed_cc_limit"] == 30.0 and p["angular_step"] == 20.0 and p["best_matches"] == 20
    assert p["mask_2d"] == [0, 0, 0, 0] and p["defocus_range"] == 500 and p["padding"] == 1
    assert p["global_search"] is False and p["local_refine"] is True and p["refine_y"] is True
    assert p["normalize"] is True and p["invert"] is False and p["threshold_rec"] is False
    p = refine3d.parse(prompts.Answers(_refine3d_heredoc(mode_local=False), "refine3d"))
    assert p["global_search"] is True and p["local_refine"] is False
    with pytest.raises(prompts.PromptError):
        refine3d.parse(prompts.Answers("a\nb\n", "refine3d"))
    bad = _refine3d_heredoc().replace("\n1.35\n", "\nnot-a-number\n")
    with pytest.raises(prompts.PromptError):
        refine3d.parse(prompts.Answers(bad, "refine3d"))


def _reconstruct_heredoc(dose=False):
    """frealign.split_reconstruction (frealign.py:1780-1824), optional dose-weighting expansion."""
    dw = ["yes", "/scratch/global_weight.txt", "yes", "4", "0.75"] if dose else ["no"]
    lines = ["/scratch/T20S_stack.mrc", "../T20S_r01_used.cistem", "null", "../T20S_r01.mrc", "T20S_r01_map1.mrc", "T20S_r01_map2.mrc",
             "output.mrc", "T20S_r01_n1.res", "C1", "1", "50", "1.35", "700.0", "0", "86.4", "2.7", "0", "2.0", "no", "0", "-1"] + dw + \
            ["0", "1", "1", "yes", "no", "no", "no", "no", "yes", "no", "no", "no", "no", "yes", "/scratch/T20S_r01_map1_n1.mrc",
             "/scratch/T20S_r01_map2_n1.mrc", "1"]
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("dose", [False, True])
def test_reconstruct3d_prompt_order(dose):
    p = reconstruct3d.parse(prompts.Answers(_reconstruct_heredoc(dose), "reconstruct3d"))
    assert p["parameters"] == "../T20S_r01_used.cistem" and p["out_statistics"] == "T20S_r01_n1.res"
    assert (p["first"], p["last"], p["symmetry"]) == (1, 50, "C1")
    assert p["outer_mask_radius"] == 86.4 and p["resolution_limit"] == 2.7 and p["score_bfactor"] == 2.0
    assert p["dose_weighting"] is dose
    if dose:
        assert p["dose_weights_file"] == "/scratch/global_weight.txt" and p["dose_fraction"] == 4 and p["dose_transition"] == 0.75
    assert p["padding"] == 1 and p["normalize"] is True and p["split_even_odd"] is True and p["dump"] is True
    assert p["dump1"].endswith("_map1_n1.mrc") and p["dump2"].endswith("_map2_n1.mrc") and p["max_threads"] == 1


def test_merge_prompt_order():
    p = local_merge3d.parse(prompts.Answers("dumpfile_map1.mrc\ndumpfile_map2.mrc\ntemp_map1_n.mrc\ntemp_map2_n.mrc\n4\n", "local_merge3d"))
    assert p == {"out1": "dumpfile_map1.mrc", "out2": "dumpfile_map2.mrc", "seed1": "temp_map1_n.mrc", "seed2": "temp_map2_n.mrc", "count": 4}
    p = merge3d.parse(prompts.Answers("a_half1.mrc\na_half2.mrc\na.mrc\na_statistics.txt\n700.0\n0\n86.4\ns/a_map1_n.mrc\ns/a_map2_n.mrc\n3\n", "merge3d"))
    assert p["filtered"] == "a.mrc" and p["molecular_mass"] == 700.0 and p["outer_radius"] == 86.4 and p["count"] == 3


def test_select_rows_and_device_pick(monkeypatch):
    from pyp_b200._lib import ROW_DTYPE

    rows = np.zeros(10, ROW_DTYPE)
    rows["position_in_stack"] = [5, 1, 9, 3, 7, 2, 10, 4, 8, 6]
    sel = refine3d.select_rows(rows, 3, 6)
    assert list(rows["position_in_stack"][sel]) == [3, 4, 5, 6]
    assert refine3d.select_rows(rows, 11, 20).size == 0
    monkeypatch.setenv("CSPB_NUM_DEVICES", "8")
    assert [prompts.pick_device(f, 100) for f in (1, 101, 201, 801)] == [0, 1, 2, 0]
    monkeypatch.setenv("CSPB_DEVICE", "3")
    assert prompts.pick_device(1, 100) == 3


def test_range_split_matches_reference_quirk():
    # local_run.py:507-516: increment = ceil(frames/cores); ranges step by increment + 1
    assert pd.split_ranges(10, 3) == [(1, 5), (6, 10)]
    assert pd.split_ranges(100, 8) == [(1, 14), (15, 28), (29, 42), (43, 56), (57, 70), (71, 84), (85, 98), (99, 100)]
    assert pd.split_ranges(0, 4) == []
    for n, w in [(10, 3), (7, 8), (100000, 8), (5, 1)]:
        shards = [pd.shard_range(1, n, r, w) for r in range(w)]
        covered = [i for lo, hi in shards for i in range(lo, hi + 1)]
        assert covered == list(range(1, n + 1))
        sizes = [hi - lo + 1 for lo, hi in shards]
        assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from pyp_b200 import dist as pd
from pyp_b200._lib import ROW_DTYPE
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
n = 11
lo, hi = pd.shard_range(1, n, rank, ws)
rows = np.zeros(hi - lo + 1, ROW_DTYPE)
rows["position_in_stack"] = np.arange(lo, hi + 1)[::-1]          # deliberately unsorted
rows["score"] = rows["position_in_stack"] * 1.5
allrows = pd.gather_rows(rows, dst=0)
acc = torch.full((64,), float(rank + 1))
pd.reduce_sum(acc, dst=0)
curve = pd.allreduce_noise_curve(np.full(5, float(rank + 1) * (hi - lo + 1)), hi - lo + 1)
if rank == 0:
    assert list(allrows["position_in_stack"]) == list(range(1, n + 1)), allrows["position_in_stack"]
    assert np.allclose(allrows["score"], np.arange(1, n + 1) * 1.5)
    assert torch.all(acc == sum(range(1, ws + 1)))
else:
    assert allrows is None
# CSP: whole particles per rank, refined entries gathered and overlaid on the input table
from pyp_b200._lib import PARTICLE_DTYPE
pinds = np.repeat(np.arange(7), 5)                      # 7 particles x 5 tilts
mine = pd.shard_entities(pinds, rank, ws)
table = np.zeros(7, PARTICLE_DTYPE); table["pind"] = np.arange(7)
part = table[np.isin(table["pind"], mine)].copy(); part["psi"] = 10.0 + part["pind"]
parts = pd.gather_table(part, dst=0)
if rank == 0:
    assert sorted(int(x) for p in parts for x in p["pind"]) == list(range(7))
    merged = pd.merge_entity_tables(table, parts, "pind")
    assert np.allclose(merged["psi"], 10.0 + np.arange(7))
want = sum((r + 1) * (pd.shard_range(1, n, r, ws)[1] - pd.shard_range(1, n, r, ws)[0] + 1) for r in range(ws)) / n
assert np.allclose(curve, want), (curve, want)
# <name>_stat.cistem over the shards of all ranks = the statistics of the merged table (particle_cspt.py:1009-1016)
from pyp_b200 import tables
rows["x_shift"] = np.sin(rows["position_in_stack"].astype(np.float64)) * 3.0
stat = pd.allreduce_parameter_statistics(rows)
full = np.zeros(n, ROW_DTYPE)
full["position_in_stack"] = np.arange(1, n + 1)
full["score"] = full["position_in_stack"] * 1.5
full["x_shift"] = np.sin(full["position_in_stack"].astype(np.float64)) * 3.0
ref = tables.parameter_statistics(full)
for name in ("score", "x_shift"):
    assert np.allclose(stat[name], ref[name], rtol=1e-5, atol=1e-5), (name, stat[name], ref[name])
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", str(script), ROOT], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_refine_ctf_prompt_order():
    """28 answers in the order of frealign.py:3998-4038."""
    from pyp_b200.cli import refine_ctf
    from pyp_b200.cli.prompts import Answers, PromptError

    a = ["s.mrc", "p.cistem", "r.mrc", "stat.txt", "no", "o.star", "c.star", "pd.mrc", "bt.mrc", "di.mrc", 1, 100, 1.35, 300.0, 0, 120.0,
         100.0, 6.0, 2000.0, "50.0", 1, "yes", "no", "yes", "no", "no", "no", "no"]
    p = refine_ctf.parse(Answers("\n".join(str(x) for x in a), "refine_ctf"))
    assert p["out_star"] == "o.star" and p["last"] == 100 and p["defocus_range"] == 2000.0 and p["refine_defocus"] and not p["beam_tilt"]
    with pytest.raises(PromptError):
        refine_ctf.parse(Answers("\n".join(str(x) for x in a[:20]), "refine_ctf"))


def test_resolution_schedule_matches_reference(tmp_path):
    """tests/golden/rhref_cases.json holds what pyp.postprocess.get_rhref (postprocess/core.py:16-55)
    returned for fixed, scheduled and FSC-driven `refine_rhref` settings."""
    import json
    import shutil

    from pyp_b200 import schedule

    g = os.path.join(ROOT, "tests", "golden")
    maps = tmp_path / "maps"
    maps.mkdir()
    shutil.copy(os.path.join(g, "rhref_fsc.txt"), maps / "ds_r01_fsc.txt")
    shutil.copy(os.path.join(g, "rhref_res.txt"), maps / "ds_r01_res.txt")
    cases = json.load(open(os.path.join(g, "rhref_cases.json")))
    assert len(cases) == 8
    for c in cases:
        got = schedule.get_rhref({"refine_rhref": c["refine_rhref"], "refine_dataset": "ds"}, c["iteration"], maps_dir=str(maps))
        assert got == pytest.approx(c["rhref"], rel=1e-12), c
    # a negative limit is jittered by at most 4 % in reciprocal space
    class R:
        @staticmethod
        def uniform(a, b):
            return b
    v = schedule.get_rhref({"refine_rhref": "-8", "refine_dataset": "ds"}, 2, rng=R)
    assert v == pytest.approx(1.0 / (1.0 / 8 + 1.0 / 8 / 25.0))
    assert schedule.get_rhref({"refine_rhref": "0", "refine_dataset": "none"}, 5, maps_dir=str(maps)) == 16


def test_refine3d_shift_prior_sources(tmp_path):
    """Answer 7 'use priors': mean / variance of the shifts come from rows 0 / 1 of the global
    `_stat.cistem` (particle_cspt.py:1009-1016) when answer 3 names one, else from the input rows."""
    from pyp_b200._lib import ROW_DTYPE
    from pyp_b200.formats import cistem

    rng = np.random.default_rng(3)
    rows = np.zeros(40, dtype=ROW_DTYPE)
    rows["position_in_stack"] = np.arange(1, 41)
    rows["x_shift"] = rng.normal(1.0, 2.0, 40)
    rows["y_shift"] = rng.normal(-0.5, 1.0, 40)
    mx, my, vx, vy = refine3d.shift_prior({"global_stat": "null"}, rows)
    assert np.isclose(mx, rows["x_shift"].mean(), atol=1e-5) and np.isclose(vy, rows["y_shift"].astype(np.float64).var(), rtol=1e-5)
    stat = np.zeros(2, dtype=ROW_DTYPE)
    stat["x_shift"] = [0.25, 9.0]
    stat["y_shift"] = [-0.75, 4.0]
    path = str(tmp_path / "ds_r01_stat.cistem")
    cistem.write_parameters(path, stat)
    assert refine3d.shift_prior({"global_stat": path}, rows) == (0.25, -0.75, 9.0, 4.0)
    assert refine3d.shift_prior({"global_stat": str(tmp_path / "missing.cistem")}, rows)[0] == mx
    assert refine3d.shift_prior({"global_stat": "null"}, rows[:0]) == (0.0, 0.0, 0.0, 0.0)


def test_likelihood_blurring_weights():
    """reconstruct3d answer 34 (frealign.py:1772,1817; legacy fan frealign.py:766-770): weights follow
    exp(LogP_k - LogP_max) with the LOGP law of SEMANTICS.md §6, members beyond the range are dropped."""
    from pyp_b200 import blur

    d = blur.offsets()
    assert d.size == 21 and d[0] == -10.0 and d[10] == 0.0 and d[-1] == 10.0
    ns = 4168
    cc = np.array([[0.30 - 0.002 * abs(k) for k in d], [0.05] * 21, [-0.1] * 21, [0.2 if k == 3 else 0.0 for k in d]])
    w = blur.weights(100 * cc, ns)
    assert np.allclose(w.sum(axis=1), 1.0) and (w >= 0).all()
    # law: ratio of two members = ((1 - cc_a^2) / (1 - cc_b^2))^(-ns/2)
    want = ((1 - cc[0, 10] ** 2) / (1 - cc[0, 9] ** 2)) ** (0.5 * ns)
    assert np.isclose(w[0, 9] / w[0, 10], want, rtol=1e-9)
    assert np.argmax(w[0]) == 10 and np.allclose(w[0], w[0][::-1])          # peaked at the refined psi, symmetric
    assert (w[0][np.abs(d) >= 8] == 0).all()                                 # > 20 LogP units below the best
    assert np.allclose(w[1], 1.0 / 21)                                        # flat likelihood: uniform fan
    assert w[2, 10] == 1.0 and w[2].sum() == 1.0                              # nothing correlates: refined pose only
    assert w[3, 13] == 1.0
    rows = np.zeros(2, dtype=ROW_DTYPE_)
    rows["psi"] = [355.0, 3.0]
    rows["theta"] = [10.0, 20.0]
    poses, idx = blur.fan_poses(rows, d)
    assert poses.shape == (42, 6) and list(idx[:21]) == [0] * 21 and poses[20, 0] == 5.0 and poses[21, 0] == 353.0
    assert (poses[:21, 1] == 10.0).all() and (poses[:, 5] == 0).all()


def test_front_ends_parse_the_reference_s_own_heredocs():
    """The stdin text is produced by the REFERENCE's command builder (frealign.py:3771-4043 mrefine_version,
    run in the build container: tests/golden/make_golden_prompts.py), not transcribed by hand."""
    import json

    from pyp_b200.cli import refine_ctf

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prompts_refine.json")))
    assert g["local"]["program"] == "refine3d" and g["beamtilt"]["program"] == "refine_ctf"
    a = prompts.Answers(g["local"]["heredoc"], "refine3d")
    p = refine3d.parse(a)
    assert a.done()                                                  # every answer consumed, none missing
    assert p["stack"] == "../T20S_stack.mrc" and p["global_stat"] == "null" and p["use_statistics"] is True and p["use_priors"] is False
    assert (p["symmetry"], p["first"], p["last"], p["percent_used"]) == ("O", 1, 100, 1)
    assert (p["pixel_size"], p["molecular_mass"], p["outer_mask_radius"]) == (1.35, 700.0, 80.0)
    assert (p["low_res_limit"], p["high_res_limit"], p["signed_cc_limit"], p["search_mask_radius"]) == (100.0, 8.0, 30.0, 120.0)
    assert p["mask_2d"] == [0, 0, 0, 0] and p["apply_2d_masking"] is False and p["defocus_range"] == 500 and p["defocus_step"] == 50.0
    assert (p["global_search"], p["local_refine"]) == (False, True)
    assert [p[f"refine_{k}"] for k in ("psi", "theta", "phi", "x", "y")] == [True] * 5
    assert p["refine_defocus"] is False and p["normalize"] is True and p["invert"] is False
    a = prompts.Answers(g["global_focus_priors"]["heredoc"], "refine3d")
    p = refine3d.parse(a)
    assert a.done() and p["use_priors"] is True and (p["global_search"], p["local_refine"]) == (True, False)
    assert p["mask_2d"] == [120.5, 98.0, 77.25, 45.0] and p["apply_2d_masking"] is True
    assert p["signed_cc_limit"] == 25.0 and p["search_mask_radius"] == 110.0 and p["defocus_range"] == 2000 and p["refine_defocus"] is True
    # refine_mask "1,0,1,1,0": the reference takes phi's flag from entry 1 (frealign.py:3814-3817)
    assert [p[f"refine_{k}"] for k in ("psi", "theta", "phi", "x", "y")] == [True, False, False, True, False]
    assert p["invert"] is True
    a = prompts.Answers(g["beamtilt"]["heredoc"], "refine_ctf")
    p = refine_ctf.parse(a)
    assert a.done() and p["beam_tilt"] is True and p["refine_defocus"] is False
    assert p["out_star"].endswith("_refined_ctf.star") and p["beamtilt_image"] == "T20S_r01_beamtilt_image.mrc"
    assert (p["first"], p["last"], p["pixel_size"], p["outer_mask_radius"], p["high_res_limit"]) == (1, 100, 1.35, 80.0, 8.0)


def test_reconstruct3d_parses_the_reference_s_own_heredocs():
    """frealign.py:1622-1835 split_reconstruction(run=False) assembled these answers
    (tests/golden/make_golden_prompts.py): plain, and dose weighting + likelihood blurring + per-particle split."""
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prompts_reconstruct.json")))
    a = prompts.Answers(g["plain"]["heredoc"], "reconstruct3d")
    p = reconstruct3d.parse(a)
    assert a.done() and g["plain"]["program"] == "reconstruct3d"
    assert p["parameters"] == "../T20S_r01_used.cistem" and p["reference"] == "../T20S_r01.mrc" and p["global_stat"] == "null"
    assert (p["symmetry"], p["first"], p["last"]) == ("O", 1, 50)
    assert (p["pixel_size"], p["outer_mask_radius"], p["resolution_limit"], p["score_bfactor"]) == (1.35, 86.4, 2.7, 2.0)
    assert p["score_weighting"] is False and p["dose_weighting"] is False and p["score_threshold"] == 0 and p["padding"] == 1
    assert p["normalize"] is True and p["split_even_odd"] is True and p["per_particle_split"] is False
    assert p["likelihood_blurring"] is False and p["dump"] is True and p["dump1"].endswith("T20S_r01_map1_n1.mrc") and p["max_threads"] == 1
    a = prompts.Answers(g["dose_blur_split"]["heredoc"], "reconstruct3d")
    p = reconstruct3d.parse(a)
    assert a.done() and p["symmetry"] == "C1"                        # reconstruct_apply_symmetry off (frealign.py:1775-1778)
    assert p["score_weighting"] is True and p["dose_weighting"] is True and p["dose_weights_file"] == "/scratch/not_provided"
    assert p["dose_multiply"] is True and p["dose_fraction"] == 4 and p["dose_transition"] == 0.75
    assert p["per_particle_split"] is True and p["likelihood_blurring"] is True
    # pyp's placeholder for "no external file" (frealign.py:1735): the weights are inferred from the parameter file itself
    rows = np.zeros(6, dtype=ROW_DTYPE_)
    rows["occupancy"], rows["tind"], rows["score"] = 100.0, [0, 0, 1, 1, 2, 2], [10, 12, 20, 22, 5, 7]
    p["pixel_size"], p["resolution_limit"] = 1.35, 2.7
    dw, note = reconstruct3d.dose_weights(p, rows, rows, 64)
    assert "the parameter file itself" in note and dw.shape == (6, 2)
    # fraction 4 of 3 indices: the best one (TIND 1) keeps the full band, the others are low-passed at 0.75 x r_rec = 0.75 x 31
    assert np.allclose(dw[:, 1], [23.25, 23.25, 0, 0, 23.25, 23.25]) and np.allclose(dw[:, 0], np.array([11, 11, 21, 21, 6, 6]) / 38 * 3)
    # without any scored projection there is nothing to weight with: said in the log, occupancies left alone
    rows["occupancy"] = 0
    dw, note = reconstruct3d.dose_weights(p, rows, rows, 64)
    assert dw is None and "skipped" in note


def test_merge_front_ends_parse_the_reference_s_own_heredocs():
    """frealign.py:1910-2136 merge_reconstructions and :1838-1903 local_merge_reconstruction, run with a recording
    executor (tests/golden/make_golden_prompts.py)."""
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prompts_merge.json")))
    assert g["merge3d"]["program"] == "merge3d" and g["local_merge3d"]["program"] == "local_merge3d"
    a = prompts.Answers(g["merge3d"]["heredoc"], "merge3d")
    p = merge3d.parse(a)
    assert a.done() and p["filtered"] == "T20S_r01_03.mrc" and p["molecular_mass"] == 700.0 and p["outer_radius"] == 86.4 and p["count"] == 3
    assert p["seed1"].endswith("T20S_r01_map1_n.mrc") and p["seed2"].endswith("T20S_r01_map2_n.mrc")
    a = prompts.Answers(g["local_merge3d"]["heredoc"], "local_merge3d")
    p = local_merge3d.parse(a)
    assert a.done() and p == {"out1": "dumpfile_map1.mrc", "out2": "dumpfile_map2.mrc", "seed1": "temp_map1_n.mrc",
                              "seed2": "temp_map2_n.mrc", "count": 3}


def test_append_stacks_drop_in(tmp_path):
    """external/cistem2/append_stacks as mrc.merge_fast drives it (src/pyp/inout/image/mrc.py:643-696): two
    answers on stdin, the first stack grows, `Error:` in the output on a dimension mismatch."""
    from pyp_b200.formats import mrc

    rng = np.random.default_rng(0)
    a, b, c = (rng.normal(size=(k, 12, 12)).astype(np.float32) for k in (3, 5, 2))
    pa, pb, pc, pd_ = (str(tmp_path / f"{n}.mrc") for n in "abcd")
    mrc.write(pa, a, 1.35)
    mrc.write(pb, b, 1.35)
    mrc.write(pc, c, 1.35)
    mrc.write(pd_, rng.normal(size=(2, 8, 8)).astype(np.float32), 1.35)
    exe = os.path.join(ROOT, "bin", "append_stacks")
    for second in (pb, pc):
        r = subprocess.run([exe], input=f"{pa}\n{second}\n", capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "Error:" not in r.stdout and "Normal termination" in r.stdout, r.stdout + r.stderr
    h, data = mrc.read(pa)
    assert (h["nx"], h["ny"], h["nz"]) == (12, 12, 10) and np.array_equal(np.asarray(data), np.concatenate([a, b, c]))
    assert os.path.getsize(pa) == 1024 + 10 * 12 * 12 * 4
    r = subprocess.run([exe], input=f"{pa}\n{pd_}\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Error:" in r.stdout            # what merge_fast looks for (mrc.py:685-686)
    assert mrc.read_header(pa)["nz"] == 10                        # untouched
    r = subprocess.run([exe], input=f"{pa}\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Error:" in r.stdout


class _FakeEngine:
    """Stand-in for pyp_b200.engine.Engine on a box without a GPU: records the calls of the front-ends and
    returns inert results, so that every host-side branch of the CLIs runs in the CPU suite (the numerics
    behind these calls are the subject of the -m gpu tests)."""
    calls = []

    def __init__(self, device=0):
        type(self).calls.append(("init", device))
        self.box = None

    @staticmethod
    def refine_defaults(box, px):
        from pyp_b200.engine import Engine_real

        return Engine_real.refine_defaults(box, px)

    @staticmethod
    def recon_defaults(box, px):
        from pyp_b200.engine import Engine_real

        return Engine_real.recon_defaults(box, px)

    def __getattr__(self, name):
        def record(*a, **k):
            type(self).calls.append((name, a, k))
            if name == "refine":
                rows = a[0].copy()
                rows["score"] = 12.5
                return rows, rows.copy(), 114 * rows.size
            if name == "band_counts":
                return 4168, 4672
            if name == "score_poses":
                idx = np.asarray(a[1])
                return np.full(idx.size, 20.0, dtype=np.float32)
            if name == "recon_get_dump":
                n = self.box * self.pad
                return np.zeros((n, n, n // 2 + 1, 4), dtype=np.float32)
            if name == "recon_finalize":
                n = self.box
                st = np.zeros((n // 2 + 1, 7), dtype=np.float32)
                st[:, 0] = np.arange(n // 2 + 1)
                return np.zeros((n, n, n), np.float32), np.zeros((n, n, n), np.float32), np.zeros((n, n, n), np.float32), st
            if name in ("refine_configure", "ensure_reference"):
                self.box = a[0].box
                if name == "ensure_reference":
                    a[2]()  # the front-end's loader reads the reference map
                    return False
            if name == "recon_begin":
                self.box, self.pad = a[0].box, a[0].pad
            return None
        return record


def test_front_end_branches_run_without_a_gpu(tmp_path, monkeypatch):
    """refine3d with priors + focus mask and reconstruct3d with likelihood blurring + dose weights, end to end
    through file I/O with the engine replaced by a recorder: the answers arrive at the engine as the C-ABI
    expects them."""
    import json

    import pyp_b200.engine as E
    from pyp_b200 import tables
    from pyp_b200.formats import cistem, mrc

    monkeypatch.setattr(E, "Engine_real", E.Engine, raising=False)
    monkeypatch.setattr(E, "Engine", _FakeEngine)
    monkeypatch.setenv("CSPB_DEVICE", "0")
    _FakeEngine.calls = []
    n, px, P = 16, 1.35, 6
    rng = np.random.default_rng(0)
    rows = np.zeros(P, dtype=ROW_DTYPE_)
    rows["position_in_stack"] = np.arange(1, P + 1)
    rows["x_shift"], rows["y_shift"] = rng.normal(0, 2, P), rng.normal(0, 1, P)
    rows["occupancy"], rows["pixel_size"], rows["score"], rows["tind"] = 100.0, px, 10.0, np.arange(P) % 3
    d = tmp_path
    mrc.write(str(d / "T20S_stack.mrc"), rng.normal(size=(P, n, n)).astype(np.float32), px)
    mrc.write(str(d / "T20S_r01.mrc"), rng.normal(size=(n, n, n)).astype(np.float32), px)
    cistem.write_parameters(str(d / "T20S_r01.cistem"), rows)
    cistem.write_parameters(str(d / "T20S_r01_stat.cistem"), tables.parameter_statistics(rows))
    (d / "statistics_r01.txt").write_text("")
    monkeypatch.chdir(d)
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prompts_refine.json")))
    text = g["global_focus_priors"]["heredoc"].replace("../T20S_stack.mrc", "T20S_stack.mrc").replace("\nnull\n", "\nT20S_r01_stat.cistem\n", 1)
    text = text.replace("\n1\n100\n", f"\n1\n{P}\n", 1).replace("\nO\n", "\nC1\n", 1)
    p = refine3d.parse(prompts.Answers(text, "refine3d"))
    import io

    log = io.StringIO()
    refine3d.run(p, out=log)
    names = [c[0] for c in _FakeEngine.calls]
    assert names.count("set_focus_mask") == 1 and "set_search_grid" in names and "refine" in names and names[-1] == "close"
    cfg = next(c for c in _FakeEngine.calls if c[0] == "ensure_reference")[1][0]
    st = tables.parameter_statistics(rows)
    assert cfg.use_priors == 1 and np.isclose(cfg.prior_mean_x, st["x_shift"][0]) and np.isclose(cfg.prior_var_y, st["y_shift"][1])
    assert (cfg.refine_psi, cfg.refine_theta, cfg.refine_phi, cfg.refine_x, cfg.refine_y) == (1, 0, 0, 1, 0) and cfg.global_search == 1
    assert next(c for c in _FakeEngine.calls if c[0] == "set_focus_mask")[1] == (120.5, 98.0, 77.25, 45.0)
    assert "Shift restraint" in log.getvalue() and "focus mask" in log.getvalue() and "Normal termination" in log.getvalue()
    out_rows = cistem.read_parameters("T20S_r01_0000001_0000100.cistem")
    assert out_rows.size == P and (out_rows["score"] == 12.5).all()
    # ---- reconstruct3d: dose weights + likelihood blurring + dumps
    _FakeEngine.calls = []
    cistem.write_parameters(str(d / "T20S_r01_used.cistem"), rows)
    tables.write_global_weights(str(d / "global_weight.txt"), tables.global_weights(rows))
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prompts_reconstruct.json")))
    text = g["dose_blur_split"]["heredoc"].replace("../T20S_stack.mrc", "T20S_stack.mrc").replace("../T20S_r01", "T20S_r01")
    text = text.replace("/scratch/not_provided", "global_weight.txt").replace("$SCRATCH/", "").replace("\n1\n50\n", f"\n1\n{P}\n", 1)
    p = reconstruct3d.parse(prompts.Answers(text, "reconstruct3d"))
    log = io.StringIO()
    reconstruct3d.run(p, out=log)
    names = [c[0] for c in _FakeEngine.calls]
    assert names.count("recon_insert") == 21 and names.count("score_poses") == 1 and "ensure_reference" in names   # the fan
    ins = [c for c in _FakeEngine.calls if c[0] == "recon_insert"]
    total_occ = sum(c[1][1]["occupancy"].astype(np.float64) for c in ins)
    dw, note = reconstruct3d.dose_weights(p, rows, rows, n)
    assert np.allclose(total_occ, rows["occupancy"], rtol=1e-5)               # flat scores: uniform fan, weights sum to 1
    assert all(np.allclose(c[1][2], dw) for c in ins) and np.allclose(dw[:, 0], 1.0)  # every fan member carries the dose pairs
    assert "Dose weighting from global_weight.txt" in log.getvalue()
    assert "Likelihood blurring" in log.getvalue() and os.path.exists("T20S_r01_map1_n1.mrc")
    # ---- local_merge3d and merge3d over that dump pair (host-side paths of both front-ends)
    _FakeEngine.calls = []
    log = io.StringIO()
    local_merge3d.run(local_merge3d.parse(prompts.Answers("sum1.mrc\nsum2.mrc\nT20S_r01_map1_n.mrc\nT20S_r01_map2_n.mrc\n1\n", "local_merge3d")), out=log)
    assert os.path.exists("sum1.mrc") and "LocalMerge3D: Normal termination" in log.getvalue()
    log = io.StringIO()
    merge3d.run(merge3d.parse(prompts.Answers("h1.mrc\nh2.mrc\nm.mrc\nst.txt\n700.0\n0\n86.4\nT20S_r01_map1_n.mrc\nT20S_r01_map2_n.mrc\n1\n", "merge3d")), out=log)
    names = [c[0] for c in _FakeEngine.calls]
    assert names.count("recon_add_dump") == 4 and "recon_finalize" in names and names[-1] == "close"
    assert os.path.exists("m.mrc") and os.path.exists("st.txt") and "Merge3D: Normal termination" in log.getvalue()


def test_tier_a_probe_and_answer_builders_match_the_reference_heredocs(tmp_path, monkeypatch):
    """oracle/tier_a.py: no real binaries in this tree -> tier B; an ELF of the right
    name and size under $PYP_DIR/external is found; the answer lists the runner feeds to the binaries equal the
    heredocs the reference's own builders produce (tests/golden/prompts_*.json)."""
    import json

    from oracle import tier_a

    GOLDEN = os.path.join(ROOT, "tests", "golden")
    monkeypatch.delenv("CSPB_TIER_A_DIR", raising=False)
    monkeypatch.setenv("PYP_DIR", "/nonexistent")
    assert tier_a.find_binaries() == {} and tier_a.probe_report()["tier"] == "B"
    ext = tmp_path / "external" / "cistem2"
    ext.mkdir(parents=True)
    (ext / "refine3d").write_bytes(b"version https://git-lfs.github.com/spec/v1\noid sha256:1865b0c3\nsize 43818480\n")  # an LFS stub
    (ext / "merge3d").write_bytes(b"\x7fELF" + b"\0" * (2 << 20))
    monkeypatch.setenv("PYP_DIR", str(tmp_path))
    rep = tier_a.probe_report()
    assert set(rep["binaries"]) == {"merge3d"} and rep["tier"] == "B" and any("refine3d" in s for s in rep["stubs"])
    g = json.load(open(os.path.join(GOLDEN, "prompts_refine.json")))
    a = tier_a.refine3d_answers("../T20S_stack.mrc", "T20S_r01.cistem", "T20S_r01.mrc", "statistics_r01.txt", "T20S_r01", 1, 100, 1.35, 700.0, 80.0,
                                "100.0", "8.0", symmetry="O", use_statistics=True, class_rhcls="8.0", search_radius=120.0, search_rhref="8.0")
    assert tier_a.heredoc(a) == g["local"]["heredoc"]
    a = tier_a.refine3d_answers("../T20S_stack.mrc", "T20S_r01.cistem", "T20S_r01.mrc", "statistics_r01.txt", "T20S_r01", 1, 100, 1.35, 700.0, 80.0,
                                "100.0", "8.0", symmetry="O", use_statistics=True, use_priors=True, signed_cc_limit="25.0", class_rhcls="8.0",
                                search_radius=110.0, search_rhref="8.0", focus=("120.5", "98.0", "77.25", "45.0"), defocus_range=2000,
                                global_search=True, local=False, mask=(1, 0, 1, 1, 0), focus_mask=True, refine_defocus=True, invert=True)
    assert tier_a.heredoc(a) == g["global_focus_priors"]["heredoc"]
    g = json.load(open(os.path.join(GOLDEN, "prompts_reconstruct.json")))
    a = tier_a.reconstruct3d_answers("../T20S_stack.mrc", "../T20S_r01_used.cistem", "../T20S_r01.mrc", "T20S_r01", 1, 50, 1.35, 700.0, 86.4, 2.7,
                                     "$SCRATCH/T20S_r01_map1_n1.mrc", "$SCRATCH/T20S_r01_map2_n1.mrc", symmetry="O")
    assert tier_a.heredoc(a) == g["plain"]["heredoc"]
    a = tier_a.reconstruct3d_answers("../T20S_stack.mrc", "../T20S_r01_used.cistem", "../T20S_r01.mrc", "T20S_r01", 1, 50, 1.35, 700.0, 86.4, 2.7,
                                     "$SCRATCH/T20S_r01_map1_n1.mrc", "$SCRATCH/T20S_r01_map2_n1.mrc", score_weighting=True,
                                     dose=("/scratch/not_provided", True, 4, 0.75), per_particle=True, blurring=True)
    assert tier_a.heredoc(a) == g["dose_blur_split"]["heredoc"]
    g = json.load(open(os.path.join(GOLDEN, "prompts_merge.json")))
    a = tier_a.merge3d_answers("T20S_r01_03", 700.0, 86.4, "$SCRATCH/T20S_r01_map1_n.mrc", "$SCRATCH/T20S_r01_map2_n.mrc", 3)
    assert tier_a.heredoc(a) == g["merge3d"]["heredoc"]
    # the range split is the reference's (local_run.py:507-516)
    assert tier_a.split_ranges(100, 8) == pd.split_ranges(100, 8)


def test_bench_workload_parameters_equal_the_library_defaults():
    """bench.py builds the workload's refine3d / reconstruct3d parameters as plain dicts so that the reference arm
    never loads libcspb200.so; they must be the library's own defaults apart from the benchmark band."""
    import bench
    from pyp_b200.engine import Engine

    for name, c in bench.CONFIGS.items():
        rp, lib_cfg = bench.refine_params(c), Engine.refine_defaults(c["box"], c["pixel"])
        differ = {k for k, v in rp.items() if abs(float(getattr(lib_cfg, k)) - float(v)) > 1e-4 * max(1.0, abs(float(v)))}
        allowed = {"mask_radius", "search_mask_radius"} | ({"global_search", "search_high_res", "search_range_x", "search_range_y"} if c.get("global_search") else set())
        assert differ <= allowed, (name, differ)
        cp, lib_rc = bench.recon_params(c), Engine.recon_defaults(c["box"], c["pixel"])
        assert {k for k, v in cp.items() if abs(float(getattr(lib_rc, k)) - float(v)) > 1e-5 * max(1.0, abs(float(v)))} <= {"pad"}, name
    # n_band of SURVEY.md §8d
    assert bench.band_count(128, 1.35, 100.0, 2.5 * 1.35) == 4168 and bench.band_count(256, 1.0, 100.0, 2.5) == 16558
    assert bench.band_count(384, 1.35, 100.0, 2.5 * 1.35) == 37174
    # reconstruct3d inserts r <= n/2 - 1 (no weight on the Nyquist planes): a few % below SURVEY's (pi/8) n^2 figure
    assert [bench.recon_band(n) for n in (128, 256, 512)] == [6227, 25309, 102135]
    # both arms print the same config object
    a = bench.parse_args(["--config", "C2"])
    b = bench.parse_args(["--config", "C2", "--impl", "reference"])
    assert bench.workload_config(a, 32768) == bench.workload_config(b, 32768)


def test_resident_server_protocol_with_a_recording_engine(tmp_path, monkeypatch):
    """pyp_b200/server.py + cli/front.py without a GPU: the daemon (here in a thread, on a recording engine) serves the
    stdlib-only client started the way bin/refine3d starts it; the files a stand-alone invocation would write appear, the
    log comes back on the client's stdout, errors come back with the word `caught` and a non-zero exit status, and the
    second request finds the reference of the first one still in place."""
    import io
    import json
    import threading

    import pyp_b200.engine as E
    from pyp_b200 import server, tables
    from pyp_b200.formats import cistem, mrc

    class _Resident(_FakeEngine):
        def ensure_reference(self, cfg, path, loader):
            key = (bytes(cfg), path)
            reused = getattr(self, "_key", None) == key
            type(self).calls.append(("ensure_reference", (cfg, path), {"reused": reused}))
            if not reused:
                loader()
            self._key, self.box = key, cfg.box
            return reused

    monkeypatch.setattr(E, "Engine_real", E.Engine, raising=False)
    monkeypatch.setattr(E, "Engine", _Resident)
    monkeypatch.setenv("CSPB_SOCKET_DIR", str(tmp_path))
    _FakeEngine.calls = []
    n, px, P = 16, 1.35, 12
    rng = np.random.default_rng(0)
    rows = np.zeros(P, dtype=ROW_DTYPE_)
    rows["position_in_stack"] = np.arange(1, P + 1)
    rows["occupancy"], rows["pixel_size"], rows["score"] = 100.0, px, 10.0
    d = tmp_path / "work"
    d.mkdir()
    mrc.write(str(d / "T20S_stack.mrc"), rng.normal(size=(P, n, n)).astype(np.float32), px)
    mrc.write(str(d / "T20S_r01.mrc"), rng.normal(size=(n, n, n)).astype(np.float32), px)
    cistem.write_parameters(str(d / "T20S_r01.cistem"), rows)
    (d / "statistics_r01.txt").write_text("")
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prompts_refine.json")))
    text = g["local"]["heredoc"].replace("../T20S_stack.mrc", "T20S_stack.mrc").replace("\nO\n", "\nC1\n", 1)
    th = threading.Thread(target=server.serve, args=(0, 0.0, 0.0), daemon=True)
    th.start()
    sock = server.socket_path(0)
    for _ in range(100):
        if os.path.exists(sock):
            break
        import time
        time.sleep(0.05)
    env = dict(os.environ, CSPB_SERVER="1", CSPB_DEVICE="0", CSPB_SOCKET_DIR=str(tmp_path), PYTHONPATH=ROOT)

    def client(stdin_text, first, last):
        t = stdin_text.replace("\n1\n100\n", f"\n{first}\n{last}\n", 1).replace("0000001_0000100", "%07d_%07d" % (first, last))
        return subprocess.run([sys.executable, "-S", "-m", "pyp_b200.cli.front", "refine3d"], input=t.encode(), cwd=str(d), env=env,
                              capture_output=True, timeout=60)

    r1 = client(text, 1, 6)
    assert r1.returncode == 0 and b"Refine3D: Normal termination" in r1.stdout, r1.stderr
    r2 = client(text, 7, 12)
    assert r2.returncode == 0 and b"Reference transform reused" in r2.stdout
    out = cistem.merge([str(d / "T20S_r01_0000001_0000006.cistem"), str(d / "T20S_r01_0000007_0000012.cistem")])
    assert list(out["position_in_stack"]) == list(range(1, P + 1)) and (out["score"] == 12.5).all()
    reuse = [c[2]["reused"] for c in _FakeEngine.calls if c[0] == "ensure_reference"]
    assert reuse == [False, True] and [c[0] for c in _FakeEngine.calls].count("init") == 1  # one context for both requests
    bad = client(text.replace("T20S_r01.cistem", "missing.cistem", 1), 1, 6)
    assert bad.returncode == 1 and b"caught" in bad.stderr
    # shutdown request
    import socket as S
    s = S.socket(S.AF_UNIX, S.SOCK_STREAM)
    s.connect(sock)
    s.sendall(b'{"prog": "shutdown"}\n')
    assert b"served 3 requests" in s.makefile("rb").readline()
    th.join(timeout=10)
    assert not th.is_alive() and not os.path.exists(sock)


def test_streamed_pipeline_batch_schedule():
    """cspb_pipeline_batches (pipeline.cu): the batches cover the stack exactly; the first one holds the 4 096 images the
    whitening curve is estimated on (or the whole stack); the middle ones are two scorer waves; the tail shrinks W, W, W/2;
    nothing smaller than half a wave except a stack that is itself smaller; a buffer never exceeds ~6 GB."""
    import ctypes as C

    from pyp_b200 import _lib

    lib = _lib.lib()
    W = 2368

    def sched(n_images, box=256, wave=W):
        buf = (C.c_int * 4096)()
        k = lib.cspb_pipeline_batches(n_images, box, wave, buf, 4096)
        assert 0 <= k <= 4096
        return list(buf[:k])

    assert sched(0) == []
    assert sched(100) == [100] and sched(4096) == [4096] and sched(5000) == [5000]
    assert sched(32768) == [4096, 3808, 4736, 4736, 4736, 4736, 2368, 2368, 1184]
    for n_images in list(range(1, 300, 7)) + [4095, 4097, 7103, 7104, 7105, 8192, 12500, 18943, 18944, 20000, 50000, 100000, 262144, 1000003]:
        s = sched(n_images)
        assert sum(s) == n_images and all(v > 0 for v in s)
        assert s[0] >= min(n_images, 4096)
        assert max(s) <= max(2 * W, 4096) + W // 2  # staging stays bounded
        if n_images >= 8 * W:
            assert s[-3:] == [W, W, W // 2]
            assert all(v == 2 * W for v in s[2:-3]) and (len(s) < 5 or s[1] == 2 * W or W // 2 <= s[1] < 2 * W)
        if len(s) > 1:
            assert min(s) >= W // 2
    # big boxes: the 6 GB staging bound wins over the wave
    for box, n_images in ((512, 20000), (1024, 3000), (384, 2048)):
        s = sched(n_images, box)
        assert sum(s) == n_images and max(s) * box * box * 4 <= (6 << 30) + W * box * box * 4
    assert lib.cspb_pipeline_batches(-1, 256, W, None, 0) < 0
