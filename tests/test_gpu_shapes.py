"""Parity of the CUDA path against the CPU oracle AT THE BASELINE SHAPES (BASELINE.json configs[1..4]):
the same comparisons as test_gpu_parity / test_gpu_csp, on the boxes, symmetry, bands and paddings the benchmark
runs — 256 px / O / 100 A..2.5 A (C2), 41 tilts x 128 px CSP (C3), 384 px global search whose reference no longer
fits the L2 (radial band order, C4) and 512 px into a 2x padded 1024^3 accumulator pair (C5).  Particle counts are
what the oracle finishes in seconds; tolerances are north_star's (score rel. err <= 1e-4 in fp32, same optimum for
>= 99.9 % of the particles, FSC >= 0.999 at every shell)."""
import gc

import numpy as np
import pytest

import bench
from common import angular_distance, pose_of
from pyp_b200 import synth, synth_torch
from pyp_b200.engine import Engine
from test_gpu_parity import SCORE_RTOL, fold_x0

pytestmark = pytest.mark.gpu


def _synthetic(cname, P, snr, seed):
    """Phantom, true rows and stack of a bench config (generated on the GPU like bench.py does — data only)."""
    import torch

    c = bench.CONFIGS[cname]
    n, px = c["box"], c["pixel"]
    dev = torch.device("cuda", 0)
    glob = c.get("global_search")
    centres, amps, sigma = synth_torch.symmetric_phantom(n, c["sym"], n_base=9 if c["sym"] != "C1" else 40,
                                                        radius_frac=0.15 if glob else 0.35, sigma=4.0 if glob else 2.0)
    vol = synth_torch.volume(n, centres, amps, sigma, dev).cpu().numpy()
    rows = synth.make_rows(P, px, seed=seed)
    stack = synth_torch.make_stack(n, centres, amps, sigma, rows, snr=snr, seed=seed + 1, device=dev).cpu().numpy()
    torch.cuda.empty_cache()
    return c, vol, rows, stack


def _same_optimum(got, want, ang_tol=2e-2, sh_tol=2e-2):
    ang = angular_distance(got, want)
    sh = np.hypot(got["x_shift"] - want["x_shift"], got["y_shift"] - want["y_shift"])
    return (ang < ang_tol) & (sh < sh_tol), ang, sh


@pytest.mark.parametrize("optimizer,evals", [(0, 18), (1, 114)])
def test_c2_256px_octahedral_refine_and_reconstruct_match_oracle(engine, oracle, optimizer, evals):
    """BASELINE configs[1]: 256-px box, 1.0 A/px, O symmetry, band 100 A..2.5 A (n_band 16 558), local search,
    reconstruction with the 24 operators (deferred on the GPU, literal in the oracle), merge3d maps."""
    from pyp_b200.symmetry import symmetry_matrices

    P = 32
    c, vol, rows, stack = _synthetic("C2", P, snr=0.05, seed=11)
    n, px = c["box"], c["pixel"]
    cfg = bench.fill(Engine.refine_defaults(n, px), bench.refine_params(c))
    cfg.optimizer = optimizer
    engine.refine_configure(cfg)
    assert engine.band_counts()[0] == 16558
    engine.set_symmetry("O")
    engine.set_reference(vol)
    engine.load_images(stack)
    ocfg = oracle.refine_cfg_from(cfg)
    curve = oracle.noise_curve(stack, ocfg)
    specs = oracle.prepare_images(stack, ocfg, curve)
    ref = oracle.Reference(vol, 1)
    # single evaluations at the true poses
    got = engine.score(rows)
    want = np.array([oracle.score(ref, specs[k], rows[k], pose_of(rows[k]), ocfg)[0] for k in range(P)])
    assert np.abs(got - want).max() <= SCORE_RTOL * np.abs(want).max()
    # the whole local refinement: 18 evaluations per particle with the analytic optimiser (coarse to fine, §7c), 114 with
    # the central-difference stencil (§7)
    start = synth.perturb_rows(rows, 2.0, 1.0)
    g, _, n_ev = engine.refine(start)
    w, n_ev_o = oracle.refine_local(ref, specs, start.astype(oracle.ROW_DTYPE), ocfg)
    assert n_ev == n_ev_o == evals * P
    same, ang, sh = _same_optimum(g, w)
    assert same.mean() >= 0.999, (np.sort(ang)[-3:], np.sort(sh)[-3:])
    # scorer parity for EVERY particle: the oracle evaluated at the pose the GPU returned reproduces the GPU's score to 1e-4
    at_g = np.array([oracle.score(ref, specs[k], g[k].astype(oracle.ROW_DTYPE), pose_of(g[k]), ocfg)[0] for k in range(P)])
    assert (np.abs(g["score"] - at_g) / np.abs(at_g)).max() <= SCORE_RTOL
    # optimiser agreement: the two optimisers stop within 0.02 deg / 0.02 A of each other; at SNR 0.05 the scores are ~5 and
    # the peak is flat, so that residual pose difference is worth up to ~1e-4 of the score (r02b: 30 of 32 below 1e-5,
    # worst 1.2e-4) — bounded at 2e-4, median at 1e-5
    rel = np.abs(g["score"] - w["score"]) / np.abs(w["score"])
    assert rel.max() <= 2 * SCORE_RTOL and np.median(rel) <= 1e-5, np.sort(rel)[-3:]
    assert np.median(angular_distance(g, rows)) < np.median(angular_distance(start, rows))
    del specs
    if optimizer == 1:  # reconstruction parity does not depend on the optimiser: checked once
        engine.set_symmetry("C1")
        return
    # reconstruct3d: insertion of the refined rows with all 24 operators, then merge3d
    rcfg = bench.fill(Engine.recon_defaults(n, px), bench.recon_params(c))
    orc = oracle.recon_cfg_from(rcfg)
    engine.recon_begin(rcfg)
    engine.recon_insert(stack, g)
    rc = oracle.Recon(orc)
    rc.insert(stack, g.astype(oracle.ROW_DTYPE), symmetry_matrices("O"))
    for h in (0, 1):
        a, b = fold_x0(engine.recon_get_dump(h)), fold_x0(rc.dump(h))
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()
        del a, b
    got_maps = engine.recon_finalize(molecular_mass_kda=440.0)
    want_maps = rc.finalize(440.0, 0.0)
    for a, b in zip(got_maps[:3], want_maps[:3]):
        f = oracle.fsc(a, b)
        assert f[1:].min() >= 0.999, f
    assert np.abs(got_maps[3][1:, 3] - want_maps[3][1:, 3]).max() < 2e-3  # FSC column of the statistics
    engine.set_symmetry("C1")
    engine.recon_end()


@pytest.mark.parametrize("mode", [5, 6])
def test_c3_41_tilts_128px_csp_matches_oracle(engine, oracle, mode):
    """BASELINE configs[2]: tilt series of 41 tilts (-60..60, 3 degree step, dose-symmetric order), 128-px box at
    1.35 A/px, per-tilt per-particle defocus, exposure window 0..20, benchmark band; particle mode 5 and micrograph mode 6."""
    from test_gpu_csp import _check

    c = bench.CONFIGS["C3"]
    n, px, n_part = c["box"], c["pixel"], 5
    ph = synth.Phantom(n, n_blobs=60, sigma=2.0)
    vol = ph.volume()
    rows, particles, tilts = synth.make_tilt_series(n_part, px, seed=3)
    assert tilts.size == 41 and rows.size == 41 * n_part
    stack = synth.make_stack(ph, rows, snr=0.2, seed=31)
    cfg = bench.fill(Engine.refine_defaults(n, px), bench.refine_params(c))
    engine.refine_configure(cfg)
    assert engine.band_counts()[0] == 4168
    engine.set_reference(vol)
    engine.load_images(stack)
    ocfg = oracle.refine_cfg_from(cfg)
    specs = oracle.prepare_images(stack, ocfg, oracle.noise_curve(stack, ocfg))
    ref = oracle.Reference(vol, 1)
    ccfg = engine.csp_defaults(mode)
    ccfg.window_min, ccfg.window_max, ccfg.iterations = 0, 20, 5
    if mode == 5:
        start_p = synth.perturb_particles(particles, 2.0, 1.5)
        start_rows = synth.rows_from_tables(rows, particles, tilts, start_p, tilts)
        got, want = _check(engine, oracle, start_rows, start_p, tilts, specs, ref, ocfg, ccfg)
    else:
        bad_t = tilts.copy()
        bad_t["shift_x"] += np.linspace(-3, 3, tilts.size).astype(np.float32)
        bad_t["angle"] += 0.5
        start_rows = synth.rows_from_tables(rows, particles, tilts, particles, bad_t)
        got, want = _check(engine, oracle, start_rows, particles, bad_t, specs, ref, ocfg, ccfg, first=0, last=11)
        assert np.array_equal(got[2][12:], bad_t[12:])


@pytest.mark.parametrize("optimizer,evals", [("analytic", 18), ("stencil", 114)])
def test_c4_384px_global_search_radial_band_matches_oracle(engine, oracle, optimizer, evals):
    """BASELINE configs[3]: 384-px box at 1.35 A/px, global search on the 20 degree grid, hits refined on the band
    100 A..2.5 px (n_band 37 174).  The half-sphere of reference quads the band touches (254 MB) exceeds the L2, so
    the band plan keeps the radial ring order (refine.cu, cspb_refine_configure) — the branch no small test reaches."""
    from pyp_b200.search_grid import search_grid

    P, K = 8, 4
    c, vol, rows, stack = _synthetic("C4", P, snr=0.1, seed=41)
    n, px = c["box"], c["pixel"]
    cfg = bench.fill(Engine.refine_defaults(n, px), bench.refine_params(c, optimizer))
    cfg.best_matches = K
    engine.refine_configure(cfg)
    assert engine.band_counts()[0] == 37174
    engine.set_symmetry("C1")
    engine.set_reference(vol)
    engine.load_images(stack)
    grid = search_grid(20.0, "C1")
    engine.set_search_grid(grid)
    ocfg = oracle.refine_cfg_from(cfg)
    specs = oracle.prepare_images(stack, ocfg, oracle.noise_curve(stack, ocfg))
    ref = oracle.Reference(vol, 1)
    # scorer parity on the radial band plan
    got = engine.score(rows)
    want = np.array([oracle.score(ref, specs[k], rows[k], pose_of(rows[k]), ocfg)[0] for k in range(P)])
    assert np.abs(got - want).max() <= SCORE_RTOL * np.abs(want).max()
    start = rows.copy()
    for k in ("psi", "theta", "phi", "x_shift", "y_shift"):
        start[k] = 0
    g, _, n_ev = engine.refine(start)
    w, n_ev_o = oracle.global_search(ref, specs, start.astype(oracle.ROW_DTYPE), ocfg, grid)
    assert n_ev == n_ev_o == P * (grid.shape[0] + K * evals)
    # discrete choices identical; the hits start up to half a grid step away, so the continuous refinement amplifies
    # fp32 summation-order noise more than a local refinement (same radius as test_global_search_matches_oracle)
    same, ang, sh = _same_optimum(g, w, 1e-1, 1e-1)
    assert same.mean() >= 0.999, (np.sort(ang)[-3:], np.sort(sh)[-3:])
    at_g = np.array([oracle.score(ref, specs[k], g[k].astype(oracle.ROW_DTYPE), pose_of(g[k]), ocfg)[0] for k in range(P)])
    assert (np.abs(g["score"] - at_g) / np.abs(at_g)).max() <= SCORE_RTOL
    rel = np.abs(g["score"] - w["score"]) / np.abs(w["score"])
    assert rel.max() <= 5 * SCORE_RTOL
    assert np.median(angular_distance(g, rows)) < 3.0  # found from scratch


def test_c5_512px_pad2_insertion_matches_oracle(engine, oracle):
    """BASELINE configs[4]: 512-px box into the 2x padded 1024^3 accumulators (2 x 8.6 GB per side), 64 particles:
    accumulator parity with the oracle half by half; the maps of that accumulator are sane (finite statistics, the
    reconstruction correlates with the phantom).  merge3d parity at padding 2 is checked on a 128-px box below —
    the oracle's plain-C 1024^3 transform would take minutes."""
    P = 64
    c, vol, rows, stack = _synthetic("C5", P, snr=0.5, seed=51)
    n, px = c["box"], c["pixel"]
    rcfg = bench.fill(Engine.recon_defaults(n, px), bench.recon_params(c))
    assert rcfg.pad == 2
    engine.set_symmetry("C1")
    engine.recon_begin(rcfg)
    engine.recon_insert(stack, rows)
    rc = oracle.Recon(oracle.recon_cfg_from(rcfg))
    rc.insert(stack, rows.astype(oracle.ROW_DTYPE))
    for h in (0, 1):
        a = engine.recon_get_dump(h)
        b = rc.dump(h)
        assert a.shape == b.shape == (1024, 1024, 513, 4)
        scale = float(np.abs(b[..., :3]).max())
        # slab by slab: 8.6 GB per array
        worst = 0.0
        for z in range(0, 1024, 64):
            worst = max(worst, float(np.abs(fold_x0_slab(a, z, z + 64) - fold_x0_slab(b, z, z + 64)).max()))
        assert worst <= 2e-5 * scale, (h, worst, scale)
        del a, b
        gc.collect()
    del rc
    gc.collect()
    m, _, _, st = engine.recon_finalize(molecular_mass_kda=800.0, want_halves=False)
    assert np.isfinite(st).all() and np.isfinite(m).all()
    # 64 projections cover Fourier space of a 512^3 map only near the origin: compare the 4x block-averaged maps
    # (128^3) on the low shells, where every voxel has been visited
    lo_m = m.reshape(128, 4, 128, 4, 128, 4).mean(axis=(1, 3, 5))
    lo_v = vol.reshape(128, 4, 128, 4, 128, 4).mean(axis=(1, 3, 5))
    f = oracle.fsc(lo_m, lo_v)
    assert f[1:9].min() > 0.7, f[:12]
    engine.recon_end()


def fold_x0_slab(d, z0, z1):
    """fold_x0 restricted to z in [z0, z1): the Friedel mate of (0, y, z) is (0, -y, -z) = indices (np - iy, np - iz)."""
    npad = d.shape[0]
    out = d[z0:z1].copy()
    for iz in range(max(z0, 1), z1):
        mate = d[npad - iz, 1:, 0, :][::-1]
        out[iz - z0, 1:, 0, 0] += mate[:, 0]
        out[iz - z0, 1:, 0, 1] -= mate[:, 1]
        out[iz - z0, 1:, 0, 2] += mate[:, 2]
    return out


def test_fold_x0_slab_equals_fold_x0():
    rng = np.random.default_rng(0)
    d = rng.normal(size=(8, 8, 5, 4)).astype(np.float32)
    full = fold_x0(d)
    got = np.concatenate([fold_x0_slab(d, 0, 3), fold_x0_slab(d, 3, 8)])
    assert np.array_equal(full, got)


def test_pad2_merge3d_maps_match_oracle_128px(engine, oracle):
    """merge3d at padding 2 (the C5 code path: 2x padded accumulators, crop + gridding correction) against the oracle
    on a 128-px box: every map FSC >= 0.999."""
    n, px, P = 128, 1.35, 96
    ph = synth.Phantom(n, n_blobs=60, sigma=2.0)
    vol = ph.volume()
    rows = synth.make_rows(P, px, seed=61)
    stack = synth.make_stack(ph, rows, snr=1.0, seed=62)
    rcfg = Engine.recon_defaults(n, px)
    rcfg.pad = 2
    engine.set_symmetry("C1")
    engine.recon_begin(rcfg)
    engine.recon_insert(stack, rows)
    got = engine.recon_finalize(molecular_mass_kda=300.0)
    rc = oracle.Recon(oracle.recon_cfg_from(rcfg))
    rc.insert(stack, rows.astype(oracle.ROW_DTYPE))
    want = rc.finalize(300.0, 0.0)
    for a, b in zip(got[:3], want[:3]):
        f = oracle.fsc(a, b)
        assert f[1:].min() >= 0.999, f
    assert np.abs(got[3][1:, 3] - want[3][1:, 3]).max() < 2e-3
    engine.recon_end()
