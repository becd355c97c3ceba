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
    rc = subprocess.run(cmd, shell=True, cwd=cwd, timeout=600).returncode
    if rc != 0:
        print(open(os.path.join(cwd, log), errors="replace").read()[-2000:])  # shown by pytest on failure
    return rc


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


def test_csp_cli_extract_and_refine(tmp_path):
    """bin/csp driven as create_csp_split_commands does (local_run.py:364-376,451-463): mode -2
    extraction from the tilt series, mode 5 over two particle ranges, mode 6 for one tilt, then the
    merge pyp performs (particle_cspt.py:95-138)."""
    from test_cpu_csp import _pose_err

    n, px, n_part = 64, 1.6, 4
    tilt_angles = np.array([-40.0, -20.0, 0.0, 20.0, 40.0])
    ph = synth.Phantom(n, n_blobs=60, sigma=1.5)
    rows, particles, tilts = synth.make_tilt_series(n_part, px, tilt_angles=tilt_angles, shift_a=3.0, extent_px=60.0, thickness_px=12.0,
                                                    defocus=22000.0)
    stack = synth.make_stack(ph, rows, snr=0.5, seed=21)
    # tilt-series images: every particle box pasted at its (ORIGINAL_X, ORIGINAL_Y) on image IMIND
    nt, ny, nx = tilts.size, 2 * n + 20, n_part * (n + 8) + 20
    series = np.zeros((nt, ny, nx), dtype=np.float32)
    for k, r in enumerate(rows):
        cx, cy = 10 + n // 2 + int(r["pind"]) * (n + 8), 10 + n // 2 + (int(r["tind"]) % 2) * n
        rows["original_x"][k], rows["original_y"][k] = cx, cy
        series[int(r["imind"]), cy - n // 2:cy + n // 2, cx - n // 2:cx + n // 2] = stack[k]
    start_p = synth.perturb_particles(particles, 2.0, 1.5)
    start_rows = synth.rows_from_tables(rows, particles, tilts, start_p, tilts)
    d = str(tmp_path)
    os.makedirs(f"{d}/frealign/maps")
    os.makedirs(f"{d}/scratch")
    par, ext = "frealign/maps/ts_r01_02.cistem", "frealign/maps/ts_r01_02_extended.cistem"
    cistem.write_parameters(f"{d}/{par}", start_rows)
    cistem.write_extended(f"{d}/{ext}", start_p, tilts)
    mrc.write(f"{d}/frealign/ts.mrc", series, px)
    mrc.write(f"{d}/scratch/ds_frames_CSP_01.mrc", ph.volume(), px)
    with open(f"{d}/.pyp_config.toml", "w") as f:
        f.write('data_set = "ds"\nextract_box = %d\nextract_bin = 1\nparticle_rad = %.2f\nrefine_rhref = "8:6.4"\nrefine_rlref = 80.0\n'
                'refine_iter = 3\ncsp_UseImagesForRefinementMin = 0\ncsp_UseImagesForRefinementMax = -1\ncsp_OptimizerMaxIter = 6\n' % (n, 0.38 * n * px))
    env = dict(os.environ, PYP_SCRATCH=f"{d}/scratch")

    def csp(*args, log="csp.log"):
        cmd = f"{BIN}/csp " + " ".join(str(a) for a in args) + f" > {log}"
        return subprocess.run(cmd, shell=True, cwd=d, env=env, timeout=600).returncode

    # mode -2 in two chunks, merged like mrc.merge_fast
    assert csp(par, ext, -2, 0, 1, 1, "frealign/ts.mrc", "frealign/ts_stack_0000_0001.mrc") == 0
    assert csp(par, ext, -2, 2, 3, 1, "frealign/ts.mrc", "frealign/ts_stack_0002_0003.mrc") == 0
    _, a = mrc.read(f"{d}/frealign/ts_stack_0000_0001.mrc")
    _, b = mrc.read(f"{d}/frealign/ts_stack_0002_0003.mrc")
    merged = np.concatenate([a, b])
    assert np.array_equal(merged, stack)
    mrc.write(f"{d}/frealign/ts_stack.mrc", merged, px)

    # mode 5 (particles) over two ranges
    for first, last in ((0, 1), (2, 3)):
        assert csp(par, ext, 5, first, last, 1, "frealign/ts.mrc", "frealign/ts_stack.mrc", log="ts_csp_%06d_%06d.log" % (first, last)) == 0
    assert "CSP: Normal termination" in open(f"{d}/ts_csp_000000_000001.log").read()
    outs = sorted(f"{d}/frealign/maps/{f}" for f in os.listdir(f"{d}/frealign/maps") if f.startswith("ts_r01_02_0") and not f.endswith("_extended.cistem"))
    assert [os.path.basename(o) for o in outs] == ["ts_r01_02_000000_000001.cistem", "ts_r01_02_000002_000003.cistem"]
    merged_rows = cistem.merge(outs)
    assert list(merged_rows["position_in_stack"]) == list(range(1, rows.size + 1))
    new_p = start_p.copy()
    for o in outs:  # dict.update semantics of Parameters.merge for the extended blocks
        pp, tt = cistem.read_extended(o.replace(".cistem", "_extended.cistem"))
        assert tt.size == 0 and pp.size == 2
        for p in pp:
            new_p[np.nonzero(new_p["pind"] == p["pind"])[0][0]] = p
    assert _pose_err(new_p, particles).mean() < 0.5 * _pose_err(start_p, particles).mean()
    assert merged_rows["score"].mean() > 0
    for p in new_p:  # PSCORE = mean score over the exposure window (update_particle_score)
        assert abs(p["score"] - merged_rows["score"][merged_rows["pind"] == p["pind"]].mean()) < 1e-3

    # mode 6 (micrographs): one tilt per process
    cistem.write_parameters(f"{d}/{par}", merged_rows)
    cistem.write_extended(f"{d}/{ext}", new_p, tilts)
    assert csp(par, ext, 6, 2, 2, 1, "frealign/ts.mrc", "frealign/ts_stack.mrc", log="/dev/null") == 0
    r6 = cistem.read_parameters(f"{d}/frealign/maps/ts_r01_02_000002_000002.cistem")
    p6, t6 = cistem.read_extended(f"{d}/frealign/maps/ts_r01_02_000002_000002_extended.cistem")
    assert p6.size == 0 and t6.size == 1 and int(t6["tind"][0]) == 2 and set(r6["tind"]) == {2} and r6.size == n_part
    assert abs(t6["angle"][0] - tilts["angle"][2]) <= 1.5 + 1e-4  # csp_ToleranceMicrographTiltAngles
    # frame (movie) refinement as pyp drives it without patches (local_run.py:434-439): mode 3, flag 0, a frame list as
    # `images`, one particle per process — the in-plane shift of every projection (frame) of the particle is refined
    frames = rows.copy()                                   # true poses; every projection displaced by a known per-frame drift
    rng = np.random.default_rng(3)
    drift = rng.uniform(-2.5, 2.5, (rows.size, 2)).astype(np.float32)
    frames["x_shift"] += drift[:, 0]
    frames["y_shift"] += drift[:, 1]
    frames["find"] = frames["tind"]
    cistem.write_parameters(f"{d}/{par}", frames)
    cistem.write_extended(f"{d}/{ext}", particles, tilts)
    open(f"{d}/frames_csp.txt", "w").write("frame_000.mrc\n")
    assert csp(par, ext, 3, 1, 1, 0, "frames_csp.txt", "frealign/ts_stack.mrc", log="frames.log") == 0
    assert "frame shifts: particles 1..1" in open(f"{d}/frames.log").read()
    rf = cistem.read_parameters(f"{d}/frealign/maps/ts_r01_02_000001_000001.cistem")
    pf, tf = cistem.read_extended(f"{d}/frealign/maps/ts_r01_02_000001_000001_extended.cistem")
    assert pf.size == 0 and tf.size == 0 and set(rf["pind"]) == {1} and rf.size == tilts.size
    sel = rows["pind"] == 1
    err0 = np.hypot(frames["x_shift"][sel] - rows["x_shift"][sel], frames["y_shift"][sel] - rows["y_shift"][sel])
    err1 = np.hypot(rf["x_shift"] - rows["x_shift"][sel], rf["y_shift"] - rows["y_shift"][sel])
    assert err1.mean() < 0.5 * err0.mean()                                           # the drift is taken out
    assert np.allclose(rf["fshift_x"], rf["x_shift"] - frames["x_shift"][sel], atol=1e-4)   # and recorded in FSHIFT_X / Y
    assert np.allclose(rf["psi"], frames["psi"][sel]) and (rf["score"] >= 0).all()
    assert csp(par, ext, -2, 0, 0, 0, "frames_csp.txt", "frealign/x.mrc", log="bad_frames.log") != 0   # extraction from frames: refused
    # an unknown mode fails loudly with pyp's failure token
    assert csp(par, ext, 9, 0, 0, 1, "frealign/ts.mrc", "frealign/ts_stack.mrc", log="bad.log") != 0
    assert "PYP (cspswarm) failed" in open(f"{d}/bad.log").read()


def test_refine_ctf_cli_recovers_defocus(tmp_path):
    """bin/refine_ctf with the answer list of frealign.py:3995-4041: a stack whose table carries a
    defocus error gets its DEFOCUS_1/2 moved back towards the truth; outputs are the star files pyp
    merges (frealign.py:3133-3154)."""
    from pyp_b200.formats import star

    from pyp_b200 import beamtilt

    n, px, n_part = 64, 1.35, 24
    ph, vol, rows, stack = small_case(n=n, n_part=n_part, snr=1.0)
    stack = beamtilt.apply_to_stack(stack, px, 300.0, 2.7, (1.5, -1.0))  # answer 23: a tilted beam to find
    start = rows.copy()
    start["defocus_1"] += 400.0
    start["defocus_2"] += 400.0
    d = str(tmp_path)
    mrc.write(f"{d}/ds_stack.mrc", stack, px)
    mrc.write(f"{d}/ds_r01.mrc", vol, px)
    cistem.write_parameters(f"{d}/ds_r01.cistem", start)
    open(f"{d}/statistics_r01.txt", "w").close()
    a = ["ds_stack.mrc", "ds_r01.cistem", "ds_r01.mrc", "statistics_r01.txt", "no", "ds_r01_0000001_0000024_refined_ctf.star",
         "ds_r01_0000001_0000024_changes.star", "ds_r01_phase_difference.mrc", "ds_r01_beamtilt_image.mrc", "ds_r01_difference_image.mrc",
         1, n_part, px, 100.0, 0, 0.38 * n * px, 60.0, 4 * px, 1000.0, "50.0", 1, "yes", "yes", "yes", "no", "no", "no", "no"]
    assert sh("refine_ctf", a, d, "ctf.log") == 0
    log = open(f"{d}/ctf.log").read()
    assert "RefineCTF: Normal termination" in log and "Beam tilt (" in log
    got = star.read_star(f"{d}/ds_r01_0000001_0000024_refined_ctf.star")
    assert got.size == n_part and list(got["position_in_stack"]) == list(range(1, n_part + 1))
    err0 = np.abs(start["defocus_1"] - rows["defocus_1"]).mean()
    err1 = np.abs(got["defocus_1"] - rows["defocus_1"]).mean()
    assert err1 < 0.6 * err0
    assert np.array_equal(got["psi"], start["psi"]) and np.allclose(got["defocus_1"] - got["defocus_2"], start["defocus_1"] - start["defocus_2"], atol=1e-2)
    chg = star.read_star(f"{d}/ds_r01_0000001_0000024_changes.star")
    assert np.allclose(chg["defocus_1"], got["defocus_1"] - start["defocus_1"], atol=1e-2)
    # beam tilt of the range (mrad) in every row, diagnostic images written (24 particles: loose bound)
    assert np.ptp(got["beam_tilt_x"]) == 0 and abs(got["beam_tilt_x"][0] - 1.5) < 0.5 and abs(got["beam_tilt_y"][0] + 1.0) < 0.5
    _, model = mrc.read(f"{d}/ds_r01_beamtilt_image.mrc")
    _, phase = mrc.read(f"{d}/ds_r01_phase_difference.mrc")
    assert model.shape[-2:] == (n, n) and np.abs(model).max() > 0.05 and np.abs(phase).max() > 0.05
