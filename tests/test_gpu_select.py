"""Score shaping and class occupancies ON THE DEVICE (select.cu) against the reference's own
outputs (tests/golden/shape_*_{in,out}.cistem, written by pyp.analysis.scores.shape_phase_residuals) and against the host
restatement pyp_b200/select.py; and the one-call refine -> select -> reconstruct pipeline that never leaves the GPU."""
import json
import os

import numpy as np
import pytest

from common import refine_cfg, small_case
from pyp_b200 import select, synth
from pyp_b200.formats import cistem

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag", ["spa", "tomo"])
def test_device_shape_scores_equal_the_reference_tables(engine, tag):
    rows = cistem.read_parameters(os.path.join(G, f"shape_{tag}_in.cistem"))
    want = cistem.read_parameters(os.path.join(G, f"shape_{tag}_out.cistem"))
    args = json.load(open(os.path.join(G, f"shape_{tag}_args.json")))
    tilts = json.load(open(os.path.join(G, f"shape_{tag}_in.json")))
    angle = select.tilt_angles_of_rows(rows, tilts)
    cfg = engine.select_defaults(args.pop("cutoff"))
    for k, v in args.items():
        setattr(cfg, k, v)
    got, thr = engine.select_scores(rows, cfg, angle)
    assert got.tobytes() == want.tobytes()  # bit for bit, every column
    assert np.isfinite(thr)


def test_device_bimodal_cutoff_with_the_host_fit(engine):
    """reconstruct_cutoff = 0: the two-Gaussian fit stays on the host (5 000-point evaluation of a scikit-learn mixture,
    statistics.py:10-148); its threshold goes to the device kernels as threshold_override."""
    pytest.importorskip("sklearn")
    rows = cistem.read_parameters(os.path.join(G, "shape_bimodal_in.cistem"))
    want = cistem.read_parameters(os.path.join(G, "shape_bimodal_out.cistem"))
    cfg = engine.select_defaults(0.0)
    cfg.threshold_override = 1.075 * select.optimal_threshold(rows["score"])
    got, thr = engine.select_scores(rows, cfg, None)
    assert got.tobytes() == want.tobytes()


def test_device_shape_scores_edges(engine):
    from pyp_b200.engine import CspbError

    rows = cistem.read_parameters(os.path.join(G, "shape_spa_in.cistem"))
    keep, _ = engine.select_scores(rows, engine.select_defaults(1.0), None)
    assert (keep["occupancy"] == 100).all()
    with pytest.raises(CspbError):
        engine.select_scores(rows, engine.select_defaults(5.0), None)   # absolute counts are not implemented
    with pytest.raises(CspbError):
        engine.select_scores(rows, engine.select_defaults(0.0), None)   # automatic cutoff needs the host fit
    assert engine.select_scores(rows[:0], engine.select_defaults(0.5), None)[0].size == 0
    # windows: against the host restatement on random settings
    rng = np.random.default_rng(2)
    for _ in range(6):
        kw = dict(mindef=float(rng.uniform(0, 15000)), maxdef=float(rng.uniform(20000, 40000)), minscore=float(rng.uniform(0, 0.2)),
                  maxscore=float(rng.uniform(0.8, 1.0)), minazh=float(rng.uniform(0, 40)), maxazh=float(rng.uniform(120, 180)))
        cut = float(rng.uniform(0.3, 1.0))
        cfg = engine.select_defaults(cut)
        for k, v in kw.items():
            setattr(cfg, k, v)
        got, _ = engine.select_scores(rows, cfg, None)
        want = select.shape_scores(rows, np.zeros(rows.size), cut, **kw)
        assert got.tobytes() == want.tobytes()


def test_device_class_occupancies(engine):
    rng = np.random.default_rng(0)
    logp = rng.normal(-2000, 3, (3, 4000)).astype(np.float32)
    sigma = rng.uniform(1, 2, (3, 4000)).astype(np.float32)
    avg = [40.0, 35.0, 25.0]
    occ, sg = engine.class_occupancies(logp, sigma, avg)
    w_occ, w_sg = select.class_occupancies(logp, sigma, avg)
    # float64 exp on both sides (CUDA's and glibc's may differ in the last place): equal after the cast to the
    # float32 OCCUPANCY column except at a rounding boundary
    assert np.abs(occ - w_occ).max() <= 2e-5 and (occ == w_occ.astype(np.float32)).mean() > 0.999
    assert np.abs(sg - w_sg).max() <= 1e-5
    assert np.allclose(occ.sum(axis=0), 100.0, atol=1e-3)


def test_device_global_weights_equal_the_reference_file(engine, tmp_path):
    """Dose weights per scan-order index on the device == tables.global_weights == the reference's global_weight.txt."""
    from pyp_b200 import tables

    rows = cistem.read_parameters(os.path.join(G, "tables_weights_in.cistem"))
    w = engine.global_weights(rows)
    assert np.array_equal(w, tables.global_weights(rows))
    out = str(tmp_path / "global_weight.txt")
    tables.write_global_weights(out, w)
    assert open(out).read() == open(os.path.join(G, "tables_global_weight.txt")).read()
    assert engine.global_weights(rows[:0]).size == 0


def test_refine_select_reconstruct_in_one_call(engine):
    """cspb_refine_select_reconstruct == refine3d -> shape_phase_residuals on the host -> reconstruct3d."""
    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=160, snr=0.3)
    start = synth.perturb_rows(rows, 2.0, 1.0)
    cfg = refine_cfg(n, px)
    rc = engine.recon_defaults(n, px)
    engine.set_symmetry("C1")
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    engine.load_images(stack)
    refined, _, n_ev = engine.refine(start)
    shaped = select.shape_scores(refined, np.zeros(refined.size), 0.7)
    engine.recon_begin(rc)
    engine.recon_insert(stack, shaped)
    want = [engine.recon_get_dump(h) for h in (0, 1)]
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    engine.recon_begin(rc)
    got_rows, n_ev2, thr = engine.refine_select_reconstruct(stack, start, engine.select_defaults(0.7))
    assert n_ev2 == n_ev and got_rows.tobytes() == shaped.tobytes()
    assert (got_rows["occupancy"] == 0).sum() == int(np.floor(0.3 * (refined.size - 1))) and thr == np.sort(refined["score"])[int((refined.size - 1) * 0.3)]
    for h in (0, 1):
        got = engine.recon_get_dump(h)
        assert np.abs(got - want[h]).max() <= 1e-5 * np.abs(want[h]).max()
    engine.recon_end()
