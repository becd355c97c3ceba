"""Golden fixtures for pyp_b200/tables.py: the REFERENCE's Parameters.update_particle_score and
Parameters.sync_particle_occ (src/pyp/inout/metadata/cistem_star_file.py:936-1013) applied to a small
tilt-series table.  Run in the build container only (imports /root/reference):

    python tests/golden/make_golden_tables.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401  (stubs the absent third-party modules, sets sys.path)

import numpy as np  # noqa: E402

from pyp.inout.metadata import cistem_star_file as csf  # noqa: E402


def build():
    rng = np.random.default_rng(77)
    n_part, n_tilt = 9, 7
    n = n_part * n_tilt - 5                                 # ragged: the last particle misses tilts
    data = np.zeros((n, 32), dtype=np.float64)
    data[:, 0] = np.arange(1, n + 1)
    data[:, 1:4] = rng.uniform(0, 360, (n, 3))
    data[:, 6:8] = rng.uniform(10000, 30000, (n, 2))
    data[:, 11] = rng.choice([0.0, 50.0, 100.0], n)
    data[:, 14] = rng.uniform(0, 30, n)
    data[:, 15] = 1.35
    data[:, 26] = np.arange(n) // n_tilt                    # PIND
    data[:, 27] = np.arange(n) % n_tilt                     # TIND
    data[:, 28] = 0
    keep = np.ones(n, bool)
    keep[(data[:, 26] == 4) & (data[:, 27] < 5)] = False    # particle 4 only has tilts 5, 6
    data = data[keep]
    data = data.astype(np.float32).astype(np.float64)       # what pyp holds after reading a .cistem file (float32 columns)
    particles = {p: csf.Particle(p, *rng.uniform(-5, 5, 3), *rng.uniform(0, 360, 3), *rng.uniform(0, 500, 3), 3.25, 100.0)
                 for p in range(n_part + 1)}                # particle 9 has no projection at all
    tilts = {t: {0: csf.Tilt(t, 0, 0.5 * t, -0.25 * t, -45.0 + 15.0 * t, 85.0)} for t in range(n_tilt)}
    return data, particles, tilts


def fresh(data, particles, tilts):
    import copy

    ext = csf.ExtendedParameters()
    ext.set_data(particles=copy.deepcopy(particles), tilts=copy.deepcopy(tilts))
    p = csf.Parameters()
    p.set_data(data=data.copy(), extended_parameters=ext)
    return p


def main():
    data, particles, tilts = build()
    fresh(data, particles, tilts).to_binary(os.path.join(HERE, "tables_in.cistem"))
    cases = {"tind_0_4": dict(tind_range=[0, 4]), "tind_2_open": dict(tind_range=[2, -1]),
             "angle": dict(tind_range=[], tiltang_range=[-20.0, 20.0])}
    for tag, kw in cases.items():
        p = fresh(data, particles, tilts)
        p.update_particle_score(**kw)
        p.to_binary(os.path.join(HERE, f"tables_score_{tag}.cistem"))
    p = fresh(data, particles, tilts)
    p.update_particle_score(tind_range=[0, 4])              # gives particle 4 and 9 occupancy 0
    p.sync_particle_occ()
    p.to_binary(os.path.join(HERE, "tables_sync_to_prj.cistem"))
    p = fresh(data, particles, tilts)
    p.sync_particle_occ(ptl_to_prj=False)
    p.to_binary(os.path.join(HERE, "tables_sync_to_ptl.cistem"))
    # external dose weights (inout/metadata/core.py:3039-3075) from the same table; scan-order index 3 unused
    os.environ.setdefault("PYP_DIR", "/root/reference")
    from pyp.inout.metadata import core as MC

    d2 = data.copy()
    d2[d2[:, 27] == 3, 11] = 0.0
    q = fresh(d2, particles, tilts)
    q.to_binary(os.path.join(HERE, "tables_weights_in.cistem"))
    MC.compute_global_weights(q.get_data(), os.path.join(HERE, "tables_global_weight.txt"))
    # <name>_stat.cistem exactly as particle_cspt.py:1009-1016 writes it: mean / variance of the used rows
    used = q.get_data()
    stat = csf.Parameters()
    stat.set_data(np.vstack((np.mean(used, axis=0), np.var(used, axis=0))))
    stat.to_binary(os.path.join(HERE, "tables_stat.cistem"))
    # Parameters.merge with extended files (cistem_star_file.py:656-692) as merge_alignment_parameters calls it
    # (particle_cspt.py:95-138): un-refined extended table first, then the outputs of two csp ranges
    import copy

    base = fresh(data, particles, tilts)
    base.to_binary(os.path.join(HERE, "tables_merge_base.cistem"))
    outs = []
    for tag, sel_p in (("000000_000003", range(0, 4)), ("000004_000008", range(4, 9))):
        rows_k = data[np.isin(data[:, 26], list(sel_p))].copy()
        rows_k[:, 14] += 1.0
        pk = {p: copy.deepcopy(particles[p]) for p in sel_p}
        for p in pk.values():
            p.psi += 0.5
            p.score = 20.0 + p.particle_index
        tk = {}
        if tag.startswith("000004"):                      # the second range also rewrites two tilts and adds one
            tk = {2: {0: csf.Tilt(2, 0, 9.0, 9.5, -15.25, 84.0)}, 6: {0: csf.Tilt(6, 0, 7.0, 7.5, 45.5, 86.0)},
                  9: {0: csf.Tilt(9, 0, 1.0, 2.0, 60.0, 85.0)}}
        ext = csf.ExtendedParameters()
        ext.set_data(particles=pk, tilts=tk)
        pr = csf.Parameters()
        pr.set_data(data=rows_k, extended_parameters=ext)
        path = os.path.join(HERE, f"tables_merge_{tag}.cistem")
        pr.to_binary(path)
        outs.append(path)
    merged = csf.Parameters.merge(outs[::-1], [os.path.join(HERE, "tables_merge_base_extended.cistem")] +
                                  [o.replace(".cistem", "_extended.cistem") for o in outs])
    merged.to_binary(os.path.join(HERE, "tables_merge_result.cistem"))
    print("written", sorted(f for f in os.listdir(HERE) if f.startswith("tables_")))


if __name__ == "__main__":
    main()
