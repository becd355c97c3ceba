"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (north_star): score relative error <= 1e-4 in fp32; identical best pose choices for
>= 99.9 % of particles; half-map FSC >= 0.999 at every shell.  FFT / CTF / slice building blocks
are held to fp32 round-off.
"""
import numpy as np
import pytest

from common import angular_distance, pose_of, refine_cfg, small_case

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-4


@pytest.mark.parametrize("n", [32, 64, 96, 128, 256, 384, 512])
def test_fft2_matches_oracle_and_cufft(engine, oracle, n):
    rng = np.random.default_rng(n)
    imgs = rng.normal(size=(5, n, n)).astype(np.float32)
    got = engine.fft2_r2c(imgs)
    ref = np.stack([oracle.fft2_r2c(im) for im in imgs])
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() / scale < 2e-6
    cu = engine.cufft2_r2c(imgs)
    assert np.abs(got - cu).max() / scale < 2e-6
    back = engine.fft2_c2r(got) / (n * n)
    assert np.abs(back - imgs).max() < 1e-5


def test_fft2_odd_batch_and_384(engine):
    rng = np.random.default_rng(7)
    imgs = rng.normal(size=(3, 384, 384)).astype(np.float32)
    got = engine.fft2_r2c(imgs)
    ref = np.fft.rfft2(imgs.astype(np.float64))
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6


def test_ctf_image(engine, oracle):
    _, _, rows, _ = small_case(n=64, n_part=4)
    for r in rows:
        got = engine.ctf_image(r, 64)
        ref = oracle.ctf_image(r.astype(oracle.ROW_DTYPE), 64)
        assert np.abs(got - ref).max() < 2e-3  # chi ~ 100 rad in fp32: |d chi| ~ 1e-5 * chi
    # against the float64 formula on a case with small chi error
    from pyp_b200 import synth

    r = rows[0]
    c = synth.ctf_2d(64, float(r["pixel_size"]), r["defocus_1"], r["defocus_2"], r["defocus_angle"])
    got = engine.ctf_image(r, 64)
    keep = np.ones(64, bool)
    keep[32] = False  # numpy's Nyquist row/column carry the opposite frequency sign
    assert np.abs(got[keep, :32] - c[keep, :32]).max() < 5e-3


@pytest.mark.parametrize("pad", [1, 2])
def test_projection(engine, oracle, pad):
    n, px = 64, 1.35
    _, vol, _, _ = small_case(n=n, n_part=1)
    cfg = refine_cfg(n, px, pad=pad)
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    ocfg = oracle.refine_cfg_from(cfg)
    _, r_hi = oracle.band_limits(ocfg)
    ref = oracle.Reference(vol, pad)
    for pose in [(0, 0, 0), (30, 0, 0), (0, 40, 70), (25, 60, 110), (200, 130, 300), (10, 180, 20)]:
        got = engine.project(*pose)
        want = ref.project(*pose, r_hi)
        assert np.abs(got - want).max() / np.abs(want).max() < 5e-6


def _setup(engine, oracle, n=64, n_part=24, **kw):
    px = 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=n_part)
    cfg = refine_cfg(n, px, **kw)
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    engine.load_images(stack)
    ocfg = oracle.refine_cfg_from(cfg)
    curve = oracle.noise_curve(stack, ocfg) if cfg.whiten else None
    specs = oracle.prepare_images(stack, ocfg, curve)
    ref = oracle.Reference(vol, cfg.pad)
    return ph, vol, rows, stack, cfg, ocfg, specs, ref, curve


def test_band_count(engine, oracle):
    cfg = refine_cfg(128, 1.35)
    cfg.low_res_limit, cfg.high_res_limit = 100.0, 2.5 * 1.35
    engine.refine_configure(cfg)
    n_band, n_slots = engine.band_counts()
    assert n_band == oracle.band_count(oracle.refine_cfg_from(cfg)) == 4168
    assert n_slots % 32 == 0 and n_band <= n_slots < 1.25 * n_band


def test_noise_curve(engine, oracle):
    *_, cfg, ocfg, specs, ref, curve = _setup(engine, oracle)
    got = engine.noise_curve()
    m = curve > 0
    assert np.abs(got[m] - curve[m]).max() / curve[m].max() < 1e-5


@pytest.mark.parametrize("kw", [dict(), dict(apply_mask=0), dict(whiten=0, normalize=0), dict(pad=2), dict(signed_cc_limit=0.0), dict(invert_contrast=1),
                                dict(mask_radius=0.62 * 64 * 1.35)])  # beyond the half box: normalisation radius clamped (image.py:324-331)
def test_score_matches_oracle(engine, oracle, kw):
    ph, vol, rows, stack, cfg, ocfg, specs, ref, curve = _setup(engine, oracle, **kw)
    got = engine.score(rows)
    want = np.array([oracle.score(ref, specs[k], rows[k], pose_of(rows[k]), ocfg)[0] for k in range(rows.size)])
    assert np.abs(got - want).max() <= SCORE_RTOL * np.abs(want).max()
    assert np.all(want > 5.0)  # the true poses correlate


@pytest.mark.parametrize("n", [64, 128, 256, 384, 512])  # every factorisation of the register FFT (8x8 ... 16x32)
def test_fused_preprocessing_is_bit_identical_to_the_separate_passes(engine, oracle, n, monkeypatch):
    """Once the whitening curve is known the preprocessing runs in four fused passes (fft2_whiten_mask_pack_dev); same
    butterflies in the same order as the seven separate ones, so the packed images — seen through the scores — do not
    change by a bit, and both agree with the oracle."""
    if n <= 256:
        ph, vol, rows, stack, cfg, ocfg, specs, ref, curve = _setup(engine, oracle, n=n, n_part=24 if n == 64 else 9)
    else:  # 384 = 16 x 24, 512 = 16 x 32: noise images against a noise map (a phantom of that size takes a minute on the host)
        vol, rows, stack = _noise_case(n, 5)
        cfg = refine_cfg(n, 1.35)
        engine.refine_configure(cfg)
        engine.set_reference(vol)
        engine.load_images(stack)
    first = engine.score(rows)  # first load: curve estimated on this stack, separate passes
    engine.load_images(stack)  # curve known: fused passes
    fused = engine.score(rows)
    monkeypatch.setenv("CSPB_PREP_FUSED", "0")
    engine.load_images(stack)
    separate = engine.score(rows)
    assert np.array_equal(fused, separate)
    assert np.abs(fused - first).max() <= 1e-6 * np.abs(first).max() + 1e-6
    if n <= 256:
        want = np.array([oracle.score(ref, specs[k], rows[k], pose_of(rows[k]), ocfg)[0] for k in range(rows.size)])
        assert np.abs(fused - want).max() <= SCORE_RTOL * np.abs(want).max()


def _noise_case(n, n_part, seed=7):
    import pyp_b200.synth as synth

    rng = np.random.default_rng(seed)
    vol = rng.normal(size=(n, n, n)).astype(np.float32)
    stack = rng.normal(size=(n_part, n, n)).astype(np.float32)
    return vol, synth.make_rows(n_part, 1.35, seed=seed), stack


@pytest.mark.parametrize("n,same_radius", [(64, False), (128, False), (128, True), (384, False)])
def test_kept_spectra_insertion_equals_a_second_transform(engine, oracle, n, same_radius):
    """cspb_refine_keep_spectra: the insertion rescales the forward transforms the refinement made (other normalisation
    radius: factor + DC term per image; same radius: used as they are) — accumulators equal those of its own transform."""
    import torch

    px = 1.35
    if n <= 128:
        ph, vol, rows, stack = small_case(n=n, n_part=21)
    else:
        vol, rows, stack = _noise_case(n, 21)
    cfg = refine_cfg(n, px)
    rc = engine.recon_defaults(n, px)
    if same_radius:
        rc.mask_radius = cfg.mask_radius
    engine.set_symmetry("C1")
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    dev = torch.from_numpy(stack).cuda()
    engine.keep_spectra(False)
    engine.load_images(dev)  # whitening curve estimated here (separate passes)
    engine.recon_begin(rc)
    engine.recon_insert(dev, rows)
    want = [engine.recon_get_dump(h) for h in (0, 1)]
    want_scores = engine.score(rows)
    for _ in range(2):  # curve known: fused passes, the column pass writes the plain transform beside its own output
        engine.keep_spectra(True)
        engine.load_images(dev)
        assert np.abs(engine.score(rows) - want_scores).max() <= 1e-6 * np.abs(want_scores).max()
        engine.recon_begin(rc)
        engine.recon_insert(dev[3:17], rows[3:17])  # a sub-range of the kept buffer
        engine.recon_insert(dev[:3], rows[:3])
        engine.recon_insert(dev[17:], rows[17:])
        for h in (0, 1):
            got = engine.recon_get_dump(h)
            assert np.abs(got - want[h]).max() <= 2e-6 * np.abs(want[h]).max()
    # a fresh configuration: the first load (curve unknown, separate passes) keeps its transforms too
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    engine.load_images(dev)
    engine.recon_begin(rc)
    engine.recon_insert(dev, rows)
    for h in (0, 1):
        assert np.abs(engine.recon_get_dump(h) - want[h]).max() <= 2e-6 * np.abs(want[h]).max()
    engine.keep_spectra(False)


def test_full_size_properties_o_symmetry_256(engine):
    """BASELINE configs[1] sizes (256-px box, O symmetry): properties that do not need the oracle.
    (1) a symmetric reference scores symmetry-related poses identically; (2) insertion is additive
    over disjoint particle sets; (3) projections of the phantom reconstruct it (FSC vs phantom);
    (4) refinement of the truth stays at the truth."""
    import torch

    from pyp_b200 import synth_torch
    from pyp_b200.symmetry import symmetry_matrices

    n, px, P = 256, 1.0, 192
    dev = torch.device("cuda", 0)
    centres, amps, sigma = synth_torch.symmetric_phantom(n, "O")
    vol = synth_torch.volume(n, centres, amps, sigma, dev)
    truth = __import__("pyp_b200").synth.make_rows(P, px, seed=5)
    stack = synth_torch.make_stack(n, centres, amps, sigma, truth, snr=0.2, seed=6, device=dev)
    cfg = engine.refine_defaults(n, px)
    cfg.low_res_limit, cfg.high_res_limit, cfg.mask_radius = 100.0, 2.5 * px, 0.38 * n * px
    engine.refine_configure(cfg)
    engine.set_symmetry("O")
    engine.set_reference(vol)
    engine.load_images(stack)
    s0 = engine.score(truth)
    assert (s0 > 3.0).all()
    # (1) pose M and M S^T (S in O) give the same projection of an O-symmetric map
    mats = symmetry_matrices("O").astype(np.float64)
    from pyp_b200 import csp_geometry as G
    from pyp_b200.synth import euler_matrix
    alt = truth.copy()
    for k in range(P):
        m = euler_matrix(truth["psi"][k], truth["theta"][k], truth["phi"][k]) 
        alt["psi"][k], alt["theta"][k], alt["phi"][k] = G.decode_m(mats[1 + k % 23] @ m)
    s1 = engine.score(alt)
    assert np.abs(s1 - s0).max() <= 2e-3 * np.abs(s0).max()
    # (4) refining from the truth does not walk away and never lowers the score
    refined, _, _ = engine.refine(truth)
    from common import angular_distance
    assert np.median(angular_distance(refined, truth)) < 0.3 and (refined["score"] >= s0 - 1e-3).all()
    # (2) additivity of the accumulators over disjoint sets, (3) FSC against the phantom
    rc = engine.recon_defaults(n, px)
    host = stack.cpu().numpy()
    engine.recon_begin(rc)
    engine.recon_insert(host[:96], truth[:96])
    a = [engine.recon_get_dump(h) for h in (0, 1)]
    engine.recon_begin(rc)
    engine.recon_insert(host[96:], truth[96:])
    b = [engine.recon_get_dump(h) for h in (0, 1)]
    engine.recon_begin(rc)
    engine.recon_insert(host, truth)
    for h in (0, 1):
        full = engine.recon_get_dump(h)
        assert np.abs(full - (a[h] + b[h])).max() <= 2e-5 * np.abs(full).max()
    rec, h1, h2, stats = engine.recon_finalize(molecular_mass_kda=440.0)
    v = vol.cpu().numpy()
    fa, fb = np.fft.rfftn(rec.astype(np.float64)), np.fft.rfftn(v.astype(np.float64))
    kz, ky, kx = np.meshgrid(np.fft.fftfreq(n) * n, np.fft.fftfreq(n) * n, np.fft.rfftfreq(n) * n, indexing="ij")
    shell = np.rint(np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)).astype(int)
    num = np.bincount(shell.ravel(), (fa * fb.conj()).real.ravel())
    den = np.sqrt(np.bincount(shell.ravel(), np.abs(fa).ravel() ** 2) * np.bincount(shell.ravel(), np.abs(fb).ravel() ** 2))
    fsc = num[2:32] / den[2:32]
    # 192 particles x 24 operators at SNR 0.2; the sigma = 2 px blobs carry signal to about shell 32 (8 A)
    assert fsc.min() > 0.9, fsc
    assert np.isfinite(stats).all()


def test_gather_peak_microbenchmark(engine):
    """roofline denominator of the scorer: L1/L2-resident gathers run at TB/s, an HBM-sized window slower"""
    l1 = engine.gather_peak(32 << 10, per_cta=True)
    l2 = engine.gather_peak(64 << 20)
    hbm = engine.gather_peak(2 << 30)
    assert l1 > 2000 and l2 > 2000 and hbm < l2
