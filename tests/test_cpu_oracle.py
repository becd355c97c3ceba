"""Known-answer tests of the CPU oracle itself ("parity unpinned": no reference binary or golden
vector exists for the numerics, SURVEY.md §4, so the restatement is anchored on analytic facts)."""
import numpy as np
import pytest

from common import angular_distance, pose_of, small_case
from pyp_b200 import synth


@pytest.mark.parametrize("n", [16, 48, 64, 96])
def test_fft_matches_numpy(oracle, n):
    rng = np.random.default_rng(n)
    img = rng.normal(size=(n, n)).astype(np.float32)
    got = oracle.fft2_r2c(img)
    want = np.fft.rfft2(img.astype(np.float64))
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-6
    assert np.abs(oracle.fft2_c2r(got) / (n * n) - img).max() < 1e-5


def test_ctf_matches_float64_formula(oracle):
    _, _, rows, _ = small_case(n=64, n_part=3)
    for r in rows:
        got = oracle.ctf_image(r.astype(oracle.ROW_DTYPE), 64)
        want = synth.ctf_2d(64, float(r["pixel_size"]), r["defocus_1"], r["defocus_2"], r["defocus_angle"])
        keep = np.ones(64, bool)
        keep[32] = False
        assert np.abs(got[keep, :32] - want[keep, :32]).max() < 5e-3
    assert abs(got[0, 0] - (-0.07)) < 1e-6  # CTF(0) = -amplitude contrast


def test_band_count_matches_survey(oracle):
    cfg = oracle.RefineCfg(box=128, pad=1, pixel_size=1.35, low_res_limit=100.0, high_res_limit=2.5 * 1.35)
    assert oracle.band_count(cfg) == 4168
    cfg = oracle.RefineCfg(box=256, pad=1, pixel_size=1.0, low_res_limit=100.0, high_res_limit=2.5)
    assert oracle.band_count(cfg) == 16558


@pytest.mark.parametrize("pad,tol", [(1, 0.98), (2, 0.999)])
def test_central_slice_theorem(oracle, pad, tol):
    """Fourier slice of the rendered phantom == FT of its analytic real-space projection."""
    n = 64
    ph = synth.Phantom(n, n_blobs=40, sigma=1.5)
    ref = oracle.Reference(ph.volume(), pad)
    jj, ii = np.meshgrid(np.arange(n), np.arange(n // 2 + 1), indexing="ij")
    j = np.where(jj >= n // 2, jj - n, jj)
    r = np.hypot(ii, j)
    m = (r <= 20) & (r >= 1)
    for pose in [(0, 0, 0), (0, 40, 70), (25, 60, 110), (200, 130, 300)]:
        P = ref.project(*pose, 24.0)
        F = np.fft.rfft2(ph.project(*pose)) * (-1.0) ** (ii + jj)
        cc = np.real(np.vdot(P[m], F[m])) / np.sqrt(np.vdot(P[m], P[m]).real * np.vdot(F[m], F[m]).real)
        assert cc > tol, (pose, cc)
    # the transposed convention must NOT match
    P = ref.project(25, 60, 110, 24.0)
    F = np.fft.rfft2(ph.project(110, 60, 25)) * (-1.0) ** (ii + jj)
    cc = np.real(np.vdot(P[m], F[m])) / np.sqrt(np.vdot(P[m], P[m]).real * np.vdot(F[m], F[m]).real)
    assert cc < 0.7


def _cfg(oracle, n, px, **kw):
    d = dict(box=n, pad=1, pixel_size=px, mask_radius=0.38 * n * px, low_res_limit=60.0, high_res_limit=4.0 * px, signed_cc_limit=30.0,
             defocus_step=50.0, refine_psi=1, refine_theta=1, refine_phi=1, refine_x=1, refine_y=1, refine_defocus=0, apply_mask=1,
             normalize=1, invert_contrast=0, whiten=1, local_iterations=8)
    d.update(kw)
    return oracle.RefineCfg(**d)


def test_score_peaks_at_true_pose_and_shift_sign(oracle):
    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=6, snr=1.0)
    rows = rows.astype(oracle.ROW_DTYPE)
    cfg = _cfg(oracle, n, px)
    specs = oracle.prepare_images(stack, cfg, oracle.noise_curve(stack, cfg))
    ref = oracle.Reference(vol, 1)
    for k in range(rows.size):
        p0 = np.array(pose_of(rows[k]), dtype=np.float32)
        s0, _ = oracle.score(ref, specs[k], rows[k], p0, cfg)
        assert s0 > 40
        for d in ([6, 0, 0, 0, 0, 0], [0, 6, 0, 0, 0, 0], [0, 0, 0, 3 * px, 0, 0], [0, 0, 0, 0, -3 * px, 0]):
            s1, _ = oracle.score(ref, specs[k], rows[k], p0 + np.array(d, np.float32), cfg)
            assert s1 < s0 - 1.0, (k, d, s0, s1)
        wrong = p0.copy()
        wrong[3:5] *= -1  # flipped shift sign
        if np.hypot(*p0[3:5]) > 1.5 * px:
            assert oracle.score(ref, specs[k], rows[k], wrong, cfg)[0] < s0 - 1.0


def test_analytic_gradient_matches_finite_differences(oracle):
    """SEMANTICS.md §7c: d CC / d (psi, theta, phi, x, y) from one evaluation (gradient of the trilinear interpolant,
    chain rule through the Euler matrix, derivative of the phase ramp) against central differences of the score;
    J^T J is symmetric positive definite and its shift block is the exact second moment of the projection."""
    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=6, snr=2.0)
    # signed correlation on every ring: above the signed-CC limit the gradient deliberately uses a soft sign for
    # |X_r| (continuous where the objective has kinks), which a difference quotient of the objective does not see
    cfg = _cfg(oracle, n, px, signed_cc_limit=0.0)
    specs = oracle.prepare_images(stack, cfg, oracle.noise_curve(stack, cfg))
    ref = oracle.Reference(vol, 1)
    start = synth.perturb_rows(rows, 1.0, 0.5).astype(oracle.ROW_DTYPE)

    def cc(o4):
        return o4[0] / np.sqrt(o4[2] * o4[3])

    for k in range(rows.size):
        pose = np.array(pose_of(start[k]), dtype=np.float32)
        s, o4, dnum, dB, jtj = oracle.score_grad(ref, specs[k], start[k], pose, cfg)
        s0, o40 = oracle.score(ref, specs[k], start[k], pose, cfg)
        assert abs(s - s0) <= 2e-6 * abs(s0) and np.allclose(o4, o40, rtol=2e-6)  # same sums as the plain evaluation (another interpolation order)
        g = dnum / np.sqrt(o4[2] * o4[3])
        g[:3] -= 0.5 * cc(o4) * dB / o4[3]
        for a in range(5):
            h = 5e-3
            p1, p2 = pose.copy(), pose.copy()
            p1[a] += h
            p2[a] -= h
            fd = (cc(oracle.score(ref, specs[k], start[k], p1, cfg)[1]) - cc(oracle.score(ref, specs[k], start[k], p2, cfg)[1])) / (2 * h)
            # the interpolant is piecewise trilinear: the difference quotient averages the kinks inside +-h
            assert abs(g[a] - fd) <= 0.03 * np.abs(g).max() + 2e-5, (k, a, g[a], fd)
        J = np.zeros((5, 5))
        J[np.triu_indices(5)] = jtj
        J = J + np.triu(J, 1).T
        assert np.linalg.eigvalsh(J).min() > 0


@pytest.mark.parametrize("optimizer,evals", [(0, 8 * 2 + 2), (1, 8 * (11 + 3) + 2)])
def test_local_refinement_recovers_poses(oracle, optimizer, evals):
    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=32, snr=0.5)
    cfg = _cfg(oracle, n, px, optimizer=optimizer)
    specs = oracle.prepare_images(stack, cfg, oracle.noise_curve(stack, cfg))
    ref = oracle.Reference(vol, 1)
    start = synth.perturb_rows(rows, 2.0, 1.0).astype(oracle.ROW_DTYPE)
    out, n_evals = oracle.refine_local(ref, specs, start, cfg)
    assert n_evals == rows.size * evals
    before, after = angular_distance(start, rows), angular_distance(out, rows)
    assert np.median(after) < 0.5 * np.median(before)
    sh = np.hypot(out["x_shift"] - rows["x_shift"], out["y_shift"] - rows["y_shift"]) / px
    assert np.median(sh) < 0.35
    assert (out["score"] >= 0).all() and (out["sigma"] > 0).all() and (out["logp"] < 0).all()
    # masked parameters stay untouched
    cfg2 = _cfg(oracle, n, px, refine_psi=0, refine_theta=0, refine_phi=0)
    out2, _ = oracle.refine_local(ref, specs, start, cfg2)
    assert np.allclose(out2["theta"], start["theta"]) and not np.allclose(out2["x_shift"], start["x_shift"])


def test_analytic_optimiser_reaches_the_scores_of_the_stencil_optimiser(oracle):
    """§7c against §7 on the same data: 18 evaluations per particle (coarse to fine) against 114 — the mean final score
    is not lower, no particle ends more than 1 % of its score below, and the poses are as close to the truth."""
    n, px = 96, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=48, snr=0.05)
    start = synth.perturb_rows(rows, 2.0, 1.0).astype(oracle.ROW_DTYPE)
    res = {}
    for opt in (0, 1):
        cfg = _cfg(oracle, n, px, optimizer=opt)
        cfg.low_res_limit, cfg.high_res_limit = 100.0, 2.5 * px
        specs = oracle.prepare_images(stack, cfg, oracle.noise_curve(stack, cfg))
        ref = oracle.Reference(vol, 1)
        res[opt] = oracle.refine_local(ref, specs, start, cfg)
    (lm, ev_lm), (st, ev_st) = res[0], res[1]
    assert ev_lm == 18 * rows.size and ev_st == 114 * rows.size
    assert lm["score"].mean() >= st["score"].mean() - 0.02
    assert (lm["score"] >= 0.99 * st["score"] - 0.05).all()
    assert np.median(angular_distance(lm, rows)) <= 1.1 * np.median(angular_distance(st, rows))


def test_shift_restraint_pulls_towards_the_mean(oracle):
    """refine3d answer 7 (use priors; frealign.py:3841-3844): objective = CC - (sigma^2 / N_mask) *
    sum (shift - mean)^2 / (2 var) over x, y (SEMANTICS.md §7b).  Off: the prior fields are inert."""
    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=16, snr=0.5)
    cfg = _cfg(oracle, n, px)
    specs = oracle.prepare_images(stack, cfg, oracle.noise_curve(stack, cfg))
    ref = oracle.Reference(vol, 1)
    start = synth.perturb_rows(rows, 2.0, 1.0).astype(oracle.ROW_DTYPE)
    start["sigma"] = 10.0
    free, _ = oracle.refine_local(ref, specs, start, cfg)
    inert, _ = oracle.refine_local(ref, specs, start, _cfg(oracle, n, px, use_priors=0, prior_mean_x=50.0, prior_var_x=1e-3, prior_var_y=1e-3))
    assert np.array_equal(free["x_shift"], inert["x_shift"]) and np.array_equal(free["score"], inert["score"])
    tight, n_ev = oracle.refine_local(ref, specs, start, _cfg(oracle, n, px, use_priors=1, prior_var_x=1.0, prior_var_y=1.0))
    assert n_ev == rows.size * (8 * 2 + 2)  # analytic optimiser: one gradient + one trial evaluation per iteration
    r_free = np.hypot(free["x_shift"], free["y_shift"])
    r_tight = np.hypot(tight["x_shift"], tight["y_shift"])
    assert r_tight.mean() < 0.8 * r_free.mean()
    assert (tight["score"] <= free["score"] + 1e-3).all()  # SCORE stays the unrestrained correlation
    # a non-positive variance leaves that shift free: only y is pulled
    only_y, _ = oracle.refine_local(ref, specs, start, _cfg(oracle, n, px, use_priors=1, prior_var_x=0.0, prior_var_y=1.0))
    assert np.abs(only_y["x_shift"] - free["x_shift"]).mean() < 0.25 * np.abs(tight["x_shift"] - free["x_shift"]).mean() + 1e-3
    assert np.abs(only_y["y_shift"]).mean() < 0.8 * np.abs(free["y_shift"]).mean()
    # zero sigma: the restraint vanishes
    start0 = start.copy()
    start0["sigma"] = 0.0
    a, _ = oracle.refine_local(ref, specs, start0, _cfg(oracle, n, px, use_priors=1, prior_var_x=1.0, prior_var_y=1.0))
    b, _ = oracle.refine_local(ref, specs, start0, cfg)
    assert np.array_equal(a["x_shift"], b["x_shift"])


def test_focus_mask_logp_sees_the_missing_density(oracle):
    """refine3d answers 29-32 + 44 (class_focusmask, frealign.py:3845-3848,3883-3885; SEMANTICS.md §6b):
    the images hold a blob the reference lacks; the residual inside the projected focus sphere is large
    when the sphere sits on that blob and small when it sits on the opposite side — which pins the
    projection geometry, the shift sign and the phase origin of the focus pass in one go."""
    n, px = 64, 1.35
    ph = synth.Phantom(n, n_blobs=40, sigma=1.5)
    k = int(np.argmin(np.abs(np.linalg.norm(ph.centres, axis=1) - 9.0)))  # well inside the soft mask
    ph.amps[k] = 5.0
    # small defocus and an 8 A band: the CTF delocalises the blob by lambda * defocus / d ~ 7 px, the disc is 6 px
    rows = synth.make_rows(8, px, seed=5, shift_px=3.0, defocus=(3000.0, 5000.0))
    stack = synth.make_stack(ph, rows, snr=20.0, seed=6)
    full_amp = ph.amps.copy()
    ph.amps[k] = 0.0
    vol = ph.volume()  # reference without blob k
    ph.amps = full_amp
    c = ph.centres[k]
    # geometry alone: the projected centre is where the analytic projection puts the blob
    # (no whitening here: at this SNR the noise curve is the signal spectrum and the whitened image no
    # longer resembles alpha * CTF * projection anywhere, which would bury the one missing blob)
    cfg = _cfg(oracle, n, px, whiten=0, high_res_limit=6.0 * px, focus_x=(c[0] + n // 2) * px, focus_y=(c[1] + n // 2) * px,
               focus_z=(c[2] + n // 2) * px, focus_radius=6.0 * px)
    for r in rows[:4]:
        cx, cy = oracle.focus_center(cfg, pose_of(r))
        want = c @ synth.euler_matrix(r["psi"], r["theta"], r["phi"])
        assert abs(cx - (n // 2 + want[0])) < 1e-3 and abs(cy - (n // 2 + want[1])) < 1e-3
    off = _cfg(oracle, n, px, whiten=0, high_res_limit=6.0 * px, focus_x=(-c[0] + n // 2) * px, focus_y=(-c[1] + n // 2) * px,
               focus_z=(-c[2] + n // 2) * px, focus_radius=6.0 * px)
    specs = oracle.prepare_images(stack, cfg, None)
    ref = oracle.Reference(vol, 1)
    rows = rows.astype(oracle.ROW_DTYPE)
    n_clear = 0
    for i in range(rows.size):
        pose = pose_of(rows[i])
        _, o4 = oracle.score(ref, specs[i], rows[i], np.array(pose, np.float32), cfg)
        on = oracle.focus_logp(ref, specs[i], rows[i], pose, cfg, o4)
        away = oracle.focus_logp(ref, specs[i], rows[i], pose, off, o4)
        # logp = -N/2 (1 + ln 2 pi var): same N (same radius) -> lower logp = larger residual variance
        c2 = c @ synth.euler_matrix(rows[i]["psi"], rows[i]["theta"], rows[i]["phi"])
        if 2 * np.hypot(c2[0], c2[1]) > 9.0:  # blob and mirror point are more than 1.5 disc radii apart in this view
            assert on < away - 50.0, (i, on, away)
            n_clear += 1
    assert n_clear >= 4
    # the refinement writes the masked value into LOGP, everything else unchanged
    plain, _ = oracle.refine_local(ref, specs[:2], rows[:2], _cfg(oracle, n, px, local_iterations=2))
    masked, _ = oracle.refine_local(ref, specs[:2], rows[:2], _cfg(oracle, n, px, local_iterations=2, focus_x=cfg.focus_x,
                                                                 focus_y=cfg.focus_y, focus_z=cfg.focus_z, focus_radius=cfg.focus_radius))
    assert np.array_equal(plain["score"], masked["score"]) and np.array_equal(plain["psi"], masked["psi"])
    assert not np.allclose(plain["logp"], masked["logp"])


def test_reconstruction_recovers_phantom_and_symmetry(oracle):
    from pyp_b200.symmetry import symmetry_matrices

    n, px = 32, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=240, n_blobs=20, snr=None)
    cfg = oracle.ReconCfg(box=n, pad=1, pixel_size=px, mask_radius=px * n / 2, resolution_limit=2 * px, score_bfactor=2.0, normalize=0)
    rc = oracle.Recon(cfg)
    rc.insert(stack, rows.astype(oracle.ROW_DTYPE))
    m, h1, h2, st = rc.finalize(50.0, 0.0)
    f = oracle.fsc(m, vol)
    assert f[1:8].min() > 0.95  # noiseless projections reproduce the phantom
    assert oracle.fsc(h1, h2)[1:8].min() > 0.9
    assert st.shape == (17, 7) and st[1, 1] == pytest.approx(n * px) and st[1, 3] > 0.9
    # x = 0 plane: dump keeps raw sums; the halves split by stack parity
    d0, d1 = rc.dump(0), rc.dump(1)
    assert d0[..., 2].sum() > 0 and d1[..., 2].sum() > 0 and np.all(d0[..., 3] == 0)
    # C2 symmetry doubles the accumulated weight
    rc2 = oracle.Recon(cfg)
    rc2.insert(stack[:20], rows[:20].astype(oracle.ROW_DTYPE), symmetry_matrices("C2"))
    rc1 = oracle.Recon(cfg)
    rc1.insert(stack[:20], rows[:20].astype(oracle.ROW_DTYPE))
    assert rc2.dump(0)[..., 2].sum() == pytest.approx(2 * rc1.dump(0)[..., 2].sum(), rel=1e-4)


def test_symmetry_groups_close():
    from pyp_b200.symmetry import symmetry_matrices

    for sym, order in [("C1", 1), ("C7", 7), ("D2", 4), ("D7", 14), ("T", 12), ("O", 24), ("I", 60)]:
        m = symmetry_matrices(sym).astype(np.float64)
        assert m.shape == (order, 3, 3)
        assert np.allclose(np.linalg.det(m), 1.0, atol=1e-5)
        prods = np.einsum("aij,bjk->abik", m, m).reshape(-1, 3, 3)
        for p in prods[:: max(1, order // 3)]:
            assert np.min(np.abs(m - p).reshape(order, -1).max(axis=1)) < 1e-5
    with pytest.raises(ValueError):
        symmetry_matrices("X3")


def test_beam_tilt_is_recovered_from_the_phase_sum(oracle):
    """refine_ctf answer 23 (frealign.py:3995-4041): images whose transforms carry the coma phase of a
    tilted beam -> phase sum over the particles -> least-squares fit gives the tilt back (SEMANTICS.md §12)."""
    from pyp_b200 import beamtilt

    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=48, snr=1.0)
    truth = (2.0, -1.5)  # mrad
    tilted = beamtilt.apply_to_stack(stack, px, 300.0, 2.7, truth)
    cfg = _cfg(oracle, n, px, whiten=0)
    ref = oracle.Reference(vol, 1)
    rows = rows.astype(oracle.ROW_DTYPE)
    for data, want in ((tilted, truth), (stack, (0.0, 0.0))):
        specs = oracle.prepare_images(data, cfg, None)
        S = oracle.phase_sum(ref, specs, rows, cfg)
        assert S.shape == (n, n // 2 + 1) and np.count_nonzero(S) >= oracle.band_count(cfg) - 2
        f = beamtilt.fit(S, px, 300.0, 2.7)
        assert abs(f["beam_tilt_x"] - want[0]) < 0.25 and abs(f["beam_tilt_y"] - want[1]) < 0.25, f
        assert abs(f["shift_x"]) < 0.3 and abs(f["shift_y"]) < 0.3
        assert f["phase"].shape == (n, n) and np.allclose(f["phase"][1:, 1:], -f["phase"][1:, 1:][::-1, ::-1], atol=1e-5)
    # the fit is exact on its own model
    gx = beamtilt.tilt_phase(n, px, 300.0, 2.7, (1.0, 0.5), (0.2, -0.1))
    S = np.exp(1j * gx) * (np.hypot(*np.meshgrid(np.arange(n // 2 + 1), np.fft.fftfreq(n, 1 / n))) < 24)
    f = beamtilt.fit(S, px, 300.0, 2.7)
    assert abs(f["beam_tilt_x"] - 1.0) < 1e-6 and abs(f["beam_tilt_y"] - 0.5) < 1e-6 and abs(f["shift_x"] - 0.2) < 1e-6


def test_normalisation_matches_reference_normalize_image(oracle):
    """(image - background mean) / background std over the pixels outside the particle radius,
    a radius beyond the half box clamped to it — against the reference's pyp.analysis.image.normalize_image
    (image.py:320-338,406-417; fixture by tests/golden/make_golden_normalize.py)."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "normalize_image.npz"))
    img, px = g["image"], float(g["pixel"])
    for r_a, want in zip(g["radii_angstrom"], g["normalized"]):
        got = oracle.normalize(img, r_a / px)
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max(), r_a
    # clamped radii give one and the same result; invert flips the sign; normalize = 0 is the identity
    assert np.array_equal(oracle.normalize(img, 30.0), oracle.normalize(img, 100.0))
    assert np.array_equal(oracle.normalize(img, 12.0, invert=1), -oracle.normalize(img, 12.0))
    assert np.array_equal(oracle.normalize(img, 12.0, normalize=0), img)


def test_optimised_cpu_leg_equals_the_restatement(oracle):
    """oracle/cspb_oracle_fast.c (the CPU arm bench.py times: band list, CTF once per particle, cropped reference,
    single-precision iterative FFT, lattice symmetry applied once to the sums) gives the results of cspb_oracle.c."""
    from pyp_b200.symmetry import symmetry_matrices
    from test_gpu_parity import fold_x0

    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=24, snr=0.3)
    cfg = _cfg(oracle, n, px)
    curve = oracle.noise_curve(stack, cfg)
    a = oracle.prepare_images(stack, cfg, curve)
    b = oracle.prepare_images(stack, cfg, curve, fast=True)
    assert np.abs(a - b).max() <= 3e-6 * np.abs(a).max()
    ref = oracle.Reference(vol, 1)
    start = synth.perturb_rows(rows, 2.0, 1.0).astype(oracle.ROW_DTYPE)
    r1, e1 = oracle.refine_local(ref, a, start, cfg)
    r2, e2 = oracle.refine_local_fast(ref, a, start, cfg)
    assert e1 == e2 and angular_distance(r1, r2).max() < 0.05 and np.abs(r1["score"] - r2["score"]).max() <= 2e-4 * r1["score"].max()
    rcfg = oracle.ReconCfg(box=n, pad=1, pixel_size=px, mask_radius=px * n / 2, resolution_limit=2 * px, score_bfactor=2.0, normalize=1)
    for sym in ("O", "D2", "I"):
        mats = symmetry_matrices(sym)
        x, y = oracle.Recon(rcfg), oracle.Recon(rcfg)
        x.insert(stack[:8], rows[:8].astype(oracle.ROW_DTYPE), mats)
        y.insert_fast(stack[:8], rows[:8].astype(oracle.ROW_DTYPE), mats)
        for h in (0, 1):
            d = fold_x0(x.dump(h))
            assert np.abs(d - fold_x0(y.dump(h))).max() <= 2e-5 * np.abs(d).max(), sym
