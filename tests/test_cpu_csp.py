"""CSP (external/CSP/csp) host logic and oracle on the CPU: pose composition pinned to the
reference's csp_euler_angles, known-answer refinement of a synthetic tilt series."""
import os

import numpy as np
import pytest

from pyp_b200 import csp_geometry as G
from pyp_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _lmat(psi, theta, phi):
    """pyp's left-handed matrix (geometry/core.py:176-180), z1 = phi, y = theta, z2 = psi."""
    z1, y, z2 = np.radians([phi, theta, psi])
    cz1, sz1, cy, sy, cz2, sz2 = np.cos(z1), np.sin(z1), np.cos(y), np.sin(y), np.cos(z2), np.sin(z2)
    return np.array([[cz1 * cy * cz2 - sz1 * sz2, cz1 * cy * sz2 + sz1 * cz2, -cz1 * sy],
                     [-sz1 * cy * cz2 - cz1 * sz2, -sz1 * cy * sz2 + cz1 * cz2, sz1 * sy],
                     [sy * cz2, sy * sz2, cy]])


def test_compose_matches_reference_csp_euler_angles(oracle):
    """tests/golden/csp_euler.npy was produced by the reference's csp_euler_angles
    (geometry/core.py:1081-1217): [tilt, axis, csp psi/theta/phi, translation(3), fp(5), nm(6)].
    The stored tables are particle = nm, tilt = (tilt, -axis) (inout/metadata/core.py:2923-2940)."""
    g = np.load(os.path.join(GOLD, "csp_euler.npy"))
    assert g.shape == (64, 19)
    for row in g:
        tilt, axis = row[0], row[1]
        fp, nm = row[8:13], row[13:19]
        got = G.compose_pose(nm[:3], (tilt, -axis))
        assert np.abs(_lmat(*got) - _lmat(*fp[:3])).max() < 1e-9
        assert np.abs(G.compose_shift(nm[3:6], (tilt, -axis)) - fp[3:5]).max() < 1e-9
        # the oracle's fp32 restatement of the same composition
        p = np.zeros(1, dtype=oracle.PARTICLE_DTYPE)
        p["psi"], p["theta"], p["phi"] = nm[:3]
        p["shift_x"], p["shift_y"], p["shift_z"] = nm[3:6]
        p0 = p.copy()
        p0["shift_x"] = p0["shift_y"] = p0["shift_z"] = 0.0
        t = np.zeros(1, dtype=oracle.TILT_DTYPE)
        t["angle"], t["axis"] = tilt, -axis
        o5 = oracle.csp_compose(p, p0, t, t, np.zeros(3), 1.0, (0.0, 0.0))
        assert np.abs(_lmat(*o5[:3]) - _lmat(*fp[:3])).max() < 5e-6
        assert np.abs(o5[3:5] - fp[3:5]).max() < 1e-5


def test_spa_limit_of_composition():
    """tilt 0 / axis 0: particle = minus the projection parameters (cistem_star_file.py:1363-1374)."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        psi, theta, phi = rng.uniform(0, 360), rng.uniform(1, 179), rng.uniform(0, 360)
        got = G.compose_pose((-psi, -theta, -phi), (0.0, 0.0))
        assert np.allclose(synth.euler_matrix(*got), synth.euler_matrix(psi, theta, phi), atol=1e-12)
        assert np.allclose(G.compose_shift((-1.5, 2.5, 9.0), (0.0, 0.0)), (1.5, -2.5))


def _tilt_case(oracle, n=48, n_part=4, px=2.0, snr=0.3, tilt_angles=(-40, -20, 0, 20, 40), **kw):
    ph = synth.Phantom(n, n_blobs=40, sigma=1.5)
    vol = ph.volume()
    rows, particles, tilts = synth.make_tilt_series(n_part, px, tilt_angles=np.array(tilt_angles, dtype=float), shift_a=3.0,
                                                    extent_px=40.0, thickness_px=10.0, defocus=20000.0, **kw)
    stack = synth.make_stack(ph, rows, snr=snr, seed=11)
    cfg = oracle.RefineCfg(box=n, pad=1, pixel_size=px, mask_radius=0.38 * n * px, low_res_limit=80.0, high_res_limit=4.0 * px,
                           signed_cc_limit=0.0, defocus_step=50.0, refine_psi=1, refine_theta=1, refine_phi=1, refine_x=1, refine_y=1,
                           refine_defocus=0, apply_mask=1, normalize=1, invert_contrast=0, whiten=1, local_iterations=8,
                           search_high_res=8 * px, search_range_x=10.0, search_range_y=10.0, best_matches=1, global_search=0)
    curve = oracle.noise_curve(stack, cfg)
    specs = oracle.prepare_images(stack, cfg, curve)
    ref = oracle.Reference(vol, 1)
    return rows, particles, tilts, specs, ref, cfg


def _csp_cfg(oracle, mode, **kw):
    c = oracle.CspCfg(mode=mode, window_min=0, window_max=-1, iterations=6, random_evals=0, grid_search=0, angle_step=20.0, shift_step=6.0,
                      tol_particle_psi=30.0, tol_particle_theta=30.0, tol_particle_phi=30.0, tol_particle_shift=20.0,
                      tol_tilt_angle=1.5, tol_tilt_axis=1.0, tol_tilt_shift=100.0, tol_defocus=750.0, seed=7, min_projections=0)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _pose_err(particles, truth):
    out = []
    for a, b in zip(particles, truth):
        p = synth.euler_matrix(a["psi"], a["theta"], a["phi"])
        q = synth.euler_matrix(b["psi"], b["theta"], b["phi"])
        out.append(np.degrees(np.arccos(np.clip((np.trace(p.T @ q) - 1) / 2, -1, 1))))
    return np.array(out)


def test_oracle_csp_particle_mode_recovers_truth(oracle):
    rows, particles, tilts, specs, ref, cfg = _tilt_case(oracle)
    start_p = synth.perturb_particles(particles, 2.0, 2.0)
    start_rows = synth.rows_from_tables(rows, particles, tilts, start_p, tilts)
    csp = _csp_cfg(oracle, 5)
    out_rows, out_p, out_t, n_ev = oracle.csp_run(ref, specs, start_rows, start_p, tilts, cfg, csp, 0, -1)
    nt = tilts.size
    assert n_ev == particles.size * nt * (6 * (1 + 12 + 3) + 2)
    e0, e1 = _pose_err(start_p, particles), _pose_err(out_p, particles)
    assert e1.mean() < 0.4 * e0.mean()
    s0 = np.linalg.norm([start_p[k] - particles[k] for k in ("shift_x", "shift_y", "shift_z")], axis=0)
    s1 = np.linalg.norm([out_p[k] - particles[k] for k in ("shift_x", "shift_y", "shift_z")], axis=0)
    assert s1.mean() < 0.6 * s0.mean()
    # rows are re-composed from the refined tables and the particle score is the window mean
    want = synth.rows_from_tables(start_rows, start_p, tilts, out_p, tilts)
    assert np.abs(out_rows["x_shift"] - want["x_shift"]).max() < 2e-3
    for p in out_p:
        sel = out_rows["pind"] == p["pind"]
        assert abs(p["score"] - out_rows["score"][sel].mean()) < 1e-3
    assert np.array_equal(out_t, tilts)
    # only entities first..last are touched
    _, part_p, _, _ = oracle.csp_run(ref, specs, start_rows, start_p, tilts, cfg, csp, 1, 2)
    assert np.array_equal(part_p[[0, 3]], start_p[[0, 3]]) and not np.array_equal(part_p[1], start_p[1])
    assert np.array_equal(part_p[1:3], out_p[1:3])


def test_oracle_csp_exposure_window_and_tilt_modes(oracle):
    rows, particles, tilts, specs, ref, cfg = _tilt_case(oracle, n_part=6)
    # window: only TIND 0..2 enter the objective (cistem_star_file.py:965-969)
    start_p = synth.perturb_particles(particles, 1.5, 1.0)
    start_rows = synth.rows_from_tables(rows, particles, tilts, start_p, tilts)
    csp = _csp_cfg(oracle, 5, window_max=2, iterations=2)
    out_rows, out_p, _, n_ev = oracle.csp_run(ref, specs, start_rows, start_p, tilts, cfg, csp, 0, -1)
    assert n_ev == particles.size * (3 * 2 * 16 + 2 * tilts.size)
    for p in out_p:
        sel = (out_rows["pind"] == p["pind"]) & (out_rows["tind"] <= 2)
        assert abs(p["score"] - out_rows["score"][sel].mean()) < 1e-3
    # tilt shifts (mode 3): displace one tilt image's alignment and recover it
    bad_t = tilts.copy()
    bad_t["shift_x"][1] += 4.0
    bad_t["shift_y"][1] -= 3.0
    bad_rows = synth.rows_from_tables(rows, particles, tilts, particles, bad_t)
    csp3 = _csp_cfg(oracle, 3, iterations=6)
    r3, p3, t3, _ = oracle.csp_run(ref, specs, bad_rows, particles, bad_t, cfg, csp3, 1, 1)
    assert abs(t3["shift_x"][1] - tilts["shift_x"][1]) < 0.8 and abs(t3["shift_y"][1] - tilts["shift_y"][1]) < 0.8
    assert np.array_equal(t3[[0, 2, 3, 4]], bad_t[[0, 2, 3, 4]]) and np.array_equal(p3, particles)
    # defocus offset per tilt (mode 4)
    off_rows = rows.copy()
    sel = off_rows["tind"] == 2
    off_rows["defocus_1"][sel] += 600.0
    off_rows["defocus_2"][sel] += 600.0
    csp4 = _csp_cfg(oracle, 4, iterations=6, random_evals=16)
    r4, _, _, _ = oracle.csp_run(ref, specs, off_rows, particles, tilts, cfg, csp4, 2, 2)
    assert abs(np.mean(r4["defocus_1"][sel] - rows["defocus_1"][sel])) < 200.0
    assert r4["score"][sel].mean() > 0


def test_oracle_csp_random_search_is_deterministic_and_helps(oracle):
    rows, particles, tilts, specs, ref, cfg = _tilt_case(oracle, n_part=2)
    start_p = synth.perturb_particles(particles, 8.0, 3.0, seed=9)
    start_rows = synth.rows_from_tables(rows, particles, tilts, start_p, tilts)
    csp = _csp_cfg(oracle, 5, random_evals=200, iterations=4, tol_particle_psi=15.0, tol_particle_theta=15.0,
                   tol_particle_phi=15.0, tol_particle_shift=6.0)
    a = oracle.csp_run(ref, specs, start_rows, start_p, tilts, cfg, csp, 0, -1)
    b = oracle.csp_run(ref, specs, start_rows, start_p, tilts, cfg, csp, 0, -1)
    assert np.array_equal(a[1], b[1]) and a[3] == b[3]
    assert a[3] == particles.size * tilts.size * (200 + 4 * 16 + 2)
    assert _pose_err(a[1], particles).mean() < _pose_err(start_p, particles).mean()
    with pytest.raises(ValueError):
        oracle.csp_run(ref, specs, start_rows, start_p[:1], tilts, cfg, csp, 0, -1)  # a row's particle is missing


def test_csp_cli_host_logic(tmp_path, monkeypatch):
    """argv / config / naming logic of bin/csp that needs no GPU."""
    from pyp_b200.cli import csp as cli

    assert cli.out_paths("frealign/maps/ts_r01_02.cistem", 3, 17) == (
        "frealign/maps/ts_r01_02_000003_000017.cistem", "frealign/maps/ts_r01_02_000003_000017_extended.cistem")
    assert cli.out_paths("x/ts_r01_02_region0004.cistem", 0, -1)[0] == "x/ts_r01_02_region0004_000000_-00001.cistem"
    # per-iteration colon lists: element min(iter - 2, len - 1) (project_params.py:362-373)
    assert cli.param("8:6:4", 2) == "8" and cli.param("8:6:4", 3) == "6" and cli.param("8:6:4", 9) == "4" and cli.param(5, 3) == 5
    monkeypatch.chdir(tmp_path)
    cfg = cli.load_config()
    assert cfg["csp_UseImagesForRefinementMax"] == 20 and cfg["csp_OptimizerMaxIter"] == 5  # pyp_config.toml defaults
    (tmp_path / ".pyp_config.toml").write_text('data_set = "abc"\ncsp_NumberOfRandomIterations = 50000\nrefine_iter = 4\nrefine_rhref = "8:6:4"\n')
    cfg = cli.load_config()
    assert cfg["csp_NumberOfRandomIterations"] == 50000
    monkeypatch.setenv("PYP_SCRATCH", "/scr")
    assert cli.reference_path(cfg) == "/scr/abc_frames_CSP_01.mrc"  # align/core.py:921-931
    rows = np.zeros(6, dtype=[("pind", "<i4"), ("tind", "<i4")])
    rows["pind"], rows["tind"] = [0, 0, 1, 1, 2, 2], [0, 1, 0, 1, 0, 1]
    assert list(cli.entity_rows(rows, None, None, 5, 1, 2)) == [2, 3, 4, 5]
    assert list(cli.entity_rows(rows, None, None, 6, 1, 1)) == [1, 3, 5]
    assert list(cli.entity_rows(rows, None, None, 3, 0, -1)) == [0, 1, 2, 3, 4, 5]
    # wrong argv is an error with pyp's failure token, not a crash
    import io
    out = io.StringIO()
    assert cli.main(["a", "b"], out=out) == 1 and "PYP (cspswarm) failed" in out.getvalue()


def test_defocus_offset_matches_reference():
    """Per-particle, per-tilt defocus offset (geometry/core.py:686-773 DefocusOffsetFromCenter) — closed form
    against the reference's 4x4 chain; fixture by tests/golden/make_golden_defocus.py."""
    import os

    from pyp_b200 import csp_geometry as cg

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "defocus_offset.npy"))
    for r in g:
        p, c, tilt, zo, h, want = r[0:3], r[3:6], r[6], r[13], int(r[14]), r[15]
        got = cg.defocus_offset_from_center(p, c, tilt, zo, handedness=h)
        assert abs(got - want) < 1e-9 * max(1.0, abs(want))
    # vectorised over particles; handedness is a pure sign
    P = g[:, 0:3]
    a = cg.defocus_offset_from_center(P, g[0, 3:6], 30.0, 5.0, handedness=1)
    b = cg.defocus_offset_from_center(P, g[0, 3:6], 30.0, 5.0, handedness=-1)
    assert a.shape == (g.shape[0],) and np.array_equal(a, -b)
    with pytest.raises(ValueError):
        cg.defocus_offset_from_center(P[0], g[0, 3:6], 0.0, 0.0, handedness=0)


def test_csp_cli_frame_lists(capsys):
    """Frame (movie) refinement hands `frames_csp.txt` as the images argument (local_run.py:320-323,434-439).  Extraction from a
    frame list is refused loudly with the token pyp greps for (particle_cspt.py:1663); refinement modes run on the stack
    (a missing parameter file is then the first error, not the frame list)."""
    import io

    from pyp_b200.cli import csp

    out = io.StringIO()
    rc = csp.main(["a.cistem", "a_extended.cistem", "-2", "0", "0", "0", "frames_csp.txt", "stack.mrc"], out=out)
    assert rc == 1 and "PYP (cspswarm) failed" in out.getvalue()
    assert "not implemented" in capsys.readouterr().err
    out = io.StringIO()
    rc = csp.main(["a.cistem", "a_extended.cistem", "3", "0", "0", "0", "frames_csp.txt", "stack.mrc"], out=out)
    assert rc == 1 and "a.cistem" in capsys.readouterr().err   # reached the files: the frame list itself is accepted
    rc = csp.main(["too", "few"], out=io.StringIO())
    assert rc == 1


def test_csp_argv_as_the_reference_builds_it():
    """Command lines from the reference's create_csp_split_commands (src/pyp/system/local_run.py:306-467;
    tests/golden/make_golden_prompts.py): eight arguments, driver modes 2 / 3 arrive as 5 / 6, extraction as -2
    with a per-range output stack, frame refinement as mode 3 with flag 0 and a frame list."""
    import io
    import json
    import os
    import shlex

    from pyp_b200.cli import csp

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "prompts_csp.json")))
    want = {"extract": (-2, 1, "frealign/TS_01.mrc"), "particles": (5, 1, "frealign/TS_01.mrc"),
            "micrographs": (6, 1, "frealign/TS_01.mrc"), "frames": (3, 0, "frames_csp.txt")}
    for tag, (mode, flag, images) in want.items():
        cmds = g[tag]["commands"]
        covered = []
        for c in cmds:
            argv = shlex.split(c.split(" > ")[0])[1:]
            a = csp.parse_argv(argv)
            assert (a["mode"], a["flag"], a["images"]) == (mode, flag, images)
            assert a["par"] == "TS_01_r01_02.cistem" and a["ext"] == "TS_01_r01_02_extended.cistem"
            assert a["stack"] == ("frealign/TS_01_stack_%04d_%04d.mrc" % (a["first"], a["last"]) if mode == -2 else "TS_01_stack.mrc")
            covered += list(range(a["first"], a["last"] + 1))
        n = 5 if tag == "micrographs" else 9
        assert covered == list(range(n))                       # contiguous inclusive ranges over particles / tilts
        assert cmds[0].endswith("_csp_000000_%06d.log" % csp.parse_argv(shlex.split(cmds[0].split(" > ")[0])[1:])["last"])
        assert all(c.endswith("> /dev/null") for c in cmds[1:])   # only the first range logs (local_run.py:447)
    # the frame-list command is parsed like the others (mode 3, flag 0: run_frame_shifts); here its files do not exist
    argv = shlex.split(g["frames"]["commands"][0].split(" > ")[0])[1:]
    assert csp.main(argv, out=io.StringIO()) == 1
