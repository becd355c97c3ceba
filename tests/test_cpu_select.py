"""Score shaping / occupancies against the reference's own outputs:
tests/golden/shape_*_{in,out}.cistem were produced by pyp.analysis.scores.shape_phase_residuals
(scores.py:300-761) in the build container (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from pyp_b200 import select
from pyp_b200.formats import cistem

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag", ["spa", "tomo"])
def test_shape_scores_matches_reference(tag):
    rows = cistem.read_parameters(os.path.join(G, f"shape_{tag}_in.cistem"))
    want = cistem.read_parameters(os.path.join(G, f"shape_{tag}_out.cistem"))
    args = json.load(open(os.path.join(G, f"shape_{tag}_args.json")))
    tilts = json.load(open(os.path.join(G, f"shape_{tag}_in.json")))
    angle = select.tilt_angles_of_rows(rows, tilts)
    got = select.shape_scores(rows, angle, args.pop("cutoff"), **args)
    assert np.array_equal(got["occupancy"], want["occupancy"])
    assert 0.1 < (got["occupancy"] == 0).mean() < 0.9          # the case exercises the selection
    assert got.tobytes() == want.tobytes()                      # every other column untouched, bit for bit


def test_bimodal_cutoff_matches_reference():
    """reconstruct_cutoff = 0 (scores.py:438-465): threshold = 1.075 x the crossing of the two fitted score
    populations (statistics.py:10-148).  Fixture by tests/golden/make_golden_bimodal.py from the reference's
    own functions (seed-independent there by construction)."""
    pytest.importorskip("sklearn")
    rows = cistem.read_parameters(os.path.join(G, "shape_bimodal_in.cistem"))
    want = cistem.read_parameters(os.path.join(G, "shape_bimodal_out.cistem"))
    ref = json.load(open(os.path.join(G, "shape_bimodal_threshold.json")))
    thr = select.optimal_threshold(rows["score"])
    assert abs(thr - ref["optimal_threshold"]) < 1e-6 * abs(ref["optimal_threshold"])
    for seed in (1, 7):  # the fit does not depend on its k-means start on this table
        assert abs(select.optimal_threshold(rows["score"], random_state=seed) - thr) < 1e-9
    got = select.shape_scores(rows, np.zeros(rows.size), 0.0)
    assert np.array_equal(got["occupancy"], want["occupancy"])
    assert got.tobytes() == want.tobytes()
    dropped = got["occupancy"] == 0
    assert 0.3 < dropped.mean() < 0.5 and rows["score"][dropped].max() < ref["scale"] * thr <= rows["score"][~dropped].min()
    # degenerate inputs: constant scores -> threshold 1 (statistics.py:19-21); <= 20 values -> no threshold at all
    assert select.optimal_threshold(np.full(50, 3.0)) == 1.0
    few = select.shape_scores(rows[:15], np.zeros(15), 0.0)
    assert (few["occupancy"] == 100).all()
    # one population: falls back to mean - 3 sigma of a single Gaussian
    one = np.random.default_rng(3).normal(20.0, 1.0, 2000)
    t1 = select.optimal_threshold(one)
    assert t1 < one.mean() - 2.0


def test_shape_scores_edges():
    rows = cistem.read_parameters(os.path.join(G, "shape_spa_in.cistem"))
    keep_all = select.shape_scores(rows, np.zeros(rows.size), 1.0)
    assert (keep_all["occupancy"] == 100).all()
    with pytest.raises(ValueError):
        select.shape_scores(rows, np.zeros(rows.size), 5.0)      # absolute-count cutoffs are not implemented
    with pytest.raises(ValueError):
        select.shape_scores(rows, np.zeros(3), 0.5)
    assert select.shape_scores(rows[:0], np.zeros(0), 0.5).size == 0


def test_class_occupancies_law():
    rng = np.random.default_rng(0)
    logp = rng.normal(-2000, 3, (3, 50))
    sigma = rng.uniform(1, 2, (3, 50))
    occ, sg = select.class_occupancies(logp, sigma, [40.0, 35.0, 25.0])
    assert np.allclose(occ.sum(axis=0), 100.0)
    assert np.allclose(sg, (sigma * occ / 100).sum(axis=0))
    # equal LogP -> proportional to the class averages; a class 10 or more below the best gets nothing
    occ, _ = select.class_occupancies(np.zeros((3, 4)), np.ones((3, 4)), [40.0, 35.0, 25.0])
    assert np.allclose(occ[:, 0], [40.0, 35.0, 25.0])
    lp = np.zeros((2, 1))
    lp[1] = -10.0
    occ, _ = select.class_occupancies(lp, np.ones((2, 1)), [50.0, 50.0])
    assert occ[1, 0] == 0.0 and occ[0, 0] == 100.0
    # restated line by line from occupancies.py:173-208
    K, n = logp.shape
    avg = [40.0, 35.0, 25.0]
    mx = np.amax(logp, axis=0)
    spp = np.zeros(n)
    for k in range(K):
        d = mx - logp[k]
        spp += np.where(d < 10, np.exp(-d) * avg[k], 0)
    ref = np.zeros((K, n))
    for k in range(K):
        d = mx - logp[k]
        ref[k] += np.where(d < 10, np.exp(-d) * avg[k] * 100 / spp, 0)
    assert np.allclose(select.class_occupancies(logp, sigma, avg)[0], ref)
