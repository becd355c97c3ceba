"""pyp_b200/tables.py against the reference's own Parameters.update_particle_score / sync_particle_occ
(src/pyp/inout/metadata/cistem_star_file.py:936-1013); fixtures by tests/golden/make_golden_tables.py."""
import os

import numpy as np
import pytest

from pyp_b200 import tables
from pyp_b200.formats import cistem

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(tag):
    path = os.path.join(G, f"tables_{tag}.cistem")
    rows = cistem.read_parameters(path)
    particles, tilts = cistem.read_extended(cistem.extended_path(path))
    return rows, particles, tilts


@pytest.mark.parametrize("tag,kw", [("tind_0_4", dict(tind_range=(0, 4))), ("tind_2_open", dict(tind_range=(2, -1))),
                                    ("angle", dict(tind_range=(), tiltang_range=(-20.0, 20.0)))])
def test_update_particle_score_matches_reference(tag, kw):
    rows, particles, tilts = _load("in")
    w_rows, w_particles, w_tilts = _load(f"score_{tag}")
    got = tables.update_particle_score(rows, particles, tilts, **kw)
    assert got.tobytes() == w_particles.tobytes()           # every particle column, bit for bit
    assert rows.tobytes() == w_rows.tobytes() and tilts.tobytes() == w_tilts.tobytes()
    empty = got["score"] == -1
    assert empty.any() and (got["occ"][empty] == 0).all() and (got["score"][~empty] >= 0).all()


def test_sync_particle_occ_matches_reference():
    rows, particles, tilts = _load("in")
    w_rows, w_particles, _ = _load("sync_to_prj")
    scored = tables.update_particle_score(rows, particles, tilts, tind_range=(0, 4))
    g_rows, g_particles = tables.sync_particle_occ(rows, scored)
    assert g_rows.tobytes() == w_rows.tobytes() and g_particles.tobytes() == w_particles.tobytes()
    assert (g_rows["occupancy"][g_rows["pind"] == 4] == 0).all()   # particle 4 has no tilt inside 0..4
    w_rows, w_particles, _ = _load("sync_to_ptl")
    g_rows, g_particles = tables.sync_particle_occ(rows, particles, ptl_to_prj=False)
    assert g_rows.tobytes() == w_rows.tobytes() and g_particles.tobytes() == w_particles.tobytes()


def test_edges():
    rows, particles, tilts = _load("in")
    with pytest.raises(ValueError):
        tables.update_particle_score(rows, particles, tilts, tind_range=(), tiltang_range=())
    with pytest.raises(ValueError):
        tables.update_particle_score(rows, particles, tilts, tind_range=(), tiltang_range=(10.0, -10.0))
    none = tables.update_particle_score(rows[:0], particles, tilts)
    assert (none["score"] == -1).all() and (none["occ"] == 0).all()
    r2, p2 = tables.sync_particle_occ(rows[:0], particles)
    assert r2.size == 0 and p2.tobytes() == particles.tobytes()


def test_global_weights_match_reference(tmp_path):
    """External dose weights of reconstruct3d answer 22 (inout/metadata/core.py:3039-3075): file text identical
    to the reference's compute_global_weights."""
    rows = cistem.read_parameters(os.path.join(G, "tables_weights_in.cistem"))
    w = tables.global_weights(rows)
    assert w[3] == -1.0 and (np.delete(w, 3) > 0).all()
    out = str(tmp_path / "global_weight.txt")
    tables.write_global_weights(out, w)
    assert open(out).read() == open(os.path.join(G, "tables_global_weight.txt")).read()
    assert tables.global_weights(rows[:0]).size == 0
    # the reconstruct3d front-end reads the same file
    from pyp_b200.cli import reconstruct3d

    p = {"dose_weighting": True, "dose_weights_file": out, "dose_fraction": 4, "dose_transition": 0.75, "dose_multiply": True,
         "resolution_limit": 0.0, "pixel_size": 1.0}
    dw, note = reconstruct3d.dose_weights(p, rows, rows, 64)
    assert dw.shape == (rows.size, 2) and (dw[rows["tind"] == 3] == 0).all() and out in note
    valid = w[w >= 0]
    assert np.allclose(dw[rows["tind"] == 0, 0], w[0] / valid.sum() * valid.size)      # normalised to mean 1 over the valid indices
    n_full = int(np.ceil(valid.size / 4))
    assert len({int(t) for t in rows["tind"][dw[:, 1] == 0]} - {3}) == n_full           # the best quarter keeps the whole band
    assert np.allclose(dw[(dw[:, 1] > 0), 1], 0.75 * 31.0)
    # multiply = no: the same weights normalised to sum 1
    dw2, _ = reconstruct3d.dose_weights(dict(p, dose_multiply=False), rows, rows, 64)
    assert np.allclose(dw2[:, 0] * valid.size, dw[:, 0])


def test_parameter_statistics_match_reference(tmp_path):
    """`<name>_stat.cistem` = np.mean / np.var of the used rows written as a two-row table
    (src/pyp/refine/csp/particle_cspt.py:1009-1016): byte-identical file."""
    rows = cistem.read_parameters(os.path.join(G, "tables_weights_in.cistem"))
    stat = tables.parameter_statistics(rows)
    out = str(tmp_path / "x_stat.cistem")
    cistem.write_parameters(out, stat)
    assert open(out, "rb").read() == open(os.path.join(G, "tables_stat.cistem"), "rb").read()
    # the moment form used across ranks agrees to float32 resolution
    names = rows.dtype.names
    sums = np.array([rows[n].astype(np.float64).sum() for n in names])
    sq = np.array([(rows[n].astype(np.float64) ** 2).sum() for n in names])
    mom = tables.statistics_from_moments(rows.dtype, rows.size, sums, sq)
    for n in ("x_shift", "score", "defocus_1", "psi"):
        assert np.allclose(mom[n], stat[n], rtol=1e-4, atol=1e-4)
    assert tables.parameter_statistics(rows[:0]).size == 2
    # refine3d's shift restraint reads these two rows
    from pyp_b200.cli import refine3d

    mx, my, vx, vy = refine3d.shift_prior({"global_stat": out}, rows)
    assert (mx, vx) == (float(stat["x_shift"][0]), float(stat["x_shift"][1]))


def test_merge_of_csp_outputs_matches_reference():
    """What pyp does with the files `csp` leaves behind (particle_cspt.py:95-138 -> Parameters.merge,
    cistem_star_file.py:656-692): rows stacked and sorted by POSITION_IN_STACK, extended tables overlaid on
    the un-refined one with dict.update semantics."""
    outs = [os.path.join(G, f"tables_merge_{t}.cistem") for t in ("000000_000003", "000004_000008")]
    want_rows = cistem.read_parameters(os.path.join(G, "tables_merge_result.cistem"))
    want_p, want_t = cistem.read_extended(os.path.join(G, "tables_merge_result_extended.cistem"))
    got_rows = cistem.merge(outs[::-1])
    assert got_rows.tobytes() == want_rows.tobytes()
    got_p, got_t = cistem.merge_extended([os.path.join(G, "tables_merge_base_extended.cistem")] +
                                         [cistem.extended_path(o) for o in outs])
    assert got_p.tobytes() == want_p.tobytes() and got_t.tobytes() == want_t.tobytes()
    base_p, base_t = cistem.read_extended(os.path.join(G, "tables_merge_base_extended.cistem"))
    assert got_p.size == base_p.size and got_t.size == base_t.size + 1          # one tilt added by the second range
    assert (got_p["score"][:9] == 20.0 + np.arange(9)).all() and got_p["score"][9] == base_p["score"][9]
    with pytest.raises(ValueError):
        cistem.merge_extended([])
