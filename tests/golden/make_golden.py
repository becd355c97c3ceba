"""Generate the golden fixtures in this directory by importing the REFERENCE's own Python
(/root/reference/src/pyp) in the build container.  The reference tree does not exist on the GPU
box, so the outputs are committed next to this script; re-run only when the reference changes:

    python tests/golden/make_golden.py

Companion generators (same import stubs, one topic each): make_golden_bimodal.py (automatic score cutoff),
make_golden_tables.py (particle score / occupancy bookkeeping, dose-weight file), make_golden_normalize.py
(particle normalisation), make_golden_defocus.py (per-tilt defocus offset), make_golden_prompts.py (stdin heredocs
and csp command lines assembled by the reference's own command builders).

What gets pinned:
  params_5x32.cistem / params_5x32_extended.cistem  — written by
      pyp.inout.metadata.cistem_star_file.Parameters.to_binary (cistem_star_file.py:734-776)
  volume_4x6x8.mrc, stack_3x8x8.mrc — written by pyp.inout.image.mrc.write (mrc.py:537-559)
  euler_decode.npy — (matrix -> psi,theta,phi) pairs from
      pyp.analysis.geometry.core.get_degrees_from_matrix (geometry/core.py:222-247) fed with the
      left-handed matrices that eulerTwoZYZtoOneZYZ composes (geometry/core.py:174-219)
  spa_euler.npy — spa_euler_angles inputs/outputs (geometry/core.py:250-441), the tilt/particle
      pose composition that CSP uses.
  shape_spa_{in,out}.cistem, shape_tomo_{in,out}.cistem (+ .json tilt tables) — inputs and outputs of
      pyp.analysis.scores.shape_phase_residuals (scores.py:300-761), the particle selection between
      refine3d and reconstruct3d, run with one angular / defocus group (the production default).
      Pins pyp_b200/select.py.
  rows_{a,b}.star, rows_star_merged_by_reference.npy — `.star` tables written by pyp_b200/formats/star.py and
      merged by the reference's merge_star (cistem_star_file.py:1398-1441), i.e. what pyp does with the
      outputs of refine_ctf (frealign.py:3133-3154).
  rhref_cases.json (+ rhref_fsc.txt, rhref_res.txt) — high-resolution limits returned by
      pyp.postprocess.get_rhref (postprocess/core.py:16-55) for fixed, scheduled and FSC-driven settings.
  merge_filmid_by_reference.npy — merge_all_binary_with_filmid over merge_{a,b}.cistem (cistem_star_file.py:1495-1550)
  csp_euler.npy — csp_euler_angles (geometry/core.py:1081-1217): tilt angle, axis, csp angles and
      3DAVG translation in; projection (psi, theta, phi, sx, sy) and the stored particle
      parameters (-ppsi, -ptheta, -pphi, px, py, pz) out.  Pins pyp_b200/csp_geometry.py.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"
sys.path.insert(0, REF)

# third-party modules the reference imports at module scope but which are absent here
class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})


def _ensure(name):
    try:
        __import__(name)
    except Exception:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            sub = ".".join(parts[:i])
            if sub not in sys.modules:
                sys.modules[sub] = _Stub(sub)


for _name in ["jsonrpcclient", "jsonrpcclient.requests", "jsonrpcclient.clients.http_client", "matplotlib", "matplotlib.pyplot",
              "toml", "mrcfile", "seaborn", "pymongo", "colored_traceback", "colorama", "skimage", "skimage.color", "trimesh"]:
    _ensure(_name)

import numpy as np  # noqa: E402


def main():
    from pyp.inout.metadata import cistem_star_file as csf

    rng = np.random.default_rng(42)
    n = 5
    data = np.zeros((n, 32), dtype=np.float64)
    data[:, 0] = np.arange(1, n + 1)                       # POSITION_IN_STACK
    data[:, 1:4] = rng.uniform(0, 360, (n, 3))             # PSI THETA PHI
    data[:, 4:6] = rng.uniform(-20, 20, (n, 2))            # shifts (A)
    data[:, 6:8] = rng.uniform(10000, 30000, (n, 2))       # defocus
    data[:, 8] = rng.uniform(0, 180, n)
    data[:, 9] = 0.0
    data[:, 10] = np.arange(n) % 2                         # IMAGE_IS_ACTIVE (film id)
    data[:, 11] = 100.0
    data[:, 12] = rng.uniform(-3000, -1000, n)
    data[:, 13] = rng.uniform(0.5, 3, n)
    data[:, 14] = rng.uniform(0, 30, n)
    data[:, 15] = 1.35
    data[:, 16] = 300.0
    data[:, 17] = 2.7
    data[:, 18] = 0.07
    data[:, 23:25] = rng.uniform(0, 4000, (n, 2))
    data[:, 25] = np.arange(n)                             # IMIND
    data[:, 26] = np.arange(n) // 2                        # PIND
    data[:, 27] = np.arange(n) % 3                         # TIND
    data[:, 28] = 0
    data[:, 29] = 0
    data[:, 30:32] = rng.uniform(-1, 1, (n, 2))
    p = csf.Parameters()
    particles = {}
    for pind in sorted(set(int(v) for v in data[:, 26])):
        particles[pind] = csf.Particle(pind, *rng.uniform(-5, 5, 3), *rng.uniform(0, 360, 3), *rng.uniform(0, 500, 3), 12.5, 100.0)
    tilts = {}
    for tind in sorted(set(int(v) for v in data[:, 27])):
        tilts[tind] = {0: csf.Tilt(tind, 0, 1.5 * tind, -2.0 * tind, -60.0 + 3.0 * tind, 85.3)}
    ext = csf.ExtendedParameters()
    ext.set_data(particles=particles, tilts=tilts)
    p.set_data(data=data, extended_parameters=ext)
    out = os.path.join(HERE, "params_5x32.cistem")
    p.to_binary(out)
    np.save(os.path.join(HERE, "params_5x32_data.npy"), data)
    # read back through the reference and keep what it returns (float64 view of float32 columns)
    q = csf.Parameters.from_file(out)
    np.save(os.path.join(HERE, "params_5x32_readback.npy"), q.get_data())

    # merge semantics: two shuffled halves -> sorted by POSITION_IN_STACK (cistem_star_file.py:656-692)
    a, b = csf.Parameters(), csf.Parameters()
    a.set_data(data=data[[4, 0, 2]]), b.set_data(data=data[[3, 1]])
    a.to_binary(os.path.join(HERE, "merge_a.cistem")), b.to_binary(os.path.join(HERE, "merge_b.cistem"))
    m = csf.Parameters.merge([os.path.join(HERE, "merge_a.cistem"), os.path.join(HERE, "merge_b.cistem")], [])
    m.to_binary(os.path.join(HERE, "merge_ab.cistem"))

    # per-film merge with film ids (cistem_star_file.py:1495-1550)
    mf = csf.merge_all_binary_with_filmid([os.path.join(HERE, "merge_a.cistem"), os.path.join(HERE, "merge_b.cistem")])
    np.save(os.path.join(HERE, "merge_filmid_by_reference.npy"), np.asarray(mf, dtype=np.float64))

    from pyp.inout.image import mrc

    vol = rng.normal(size=(4, 6, 8)).astype(np.float32)
    mrc.write(vol, os.path.join(HERE, "volume_4x6x8.mrc"))
    np.save(os.path.join(HERE, "volume_4x6x8.npy"), vol)
    stack = rng.normal(size=(3, 8, 8)).astype(np.float32)
    mrc.write(stack, os.path.join(HERE, "stack_3x8x8.mrc"))
    np.save(os.path.join(HERE, "stack_3x8x8.npy"), stack)

    from pyp.analysis.geometry import core as geo

    rows = []
    for _ in range(64):
        psi, theta, phi = rng.uniform(0, 360), rng.uniform(1, 179), rng.uniform(0, 360)
        # left-handed matrix of geometry/core.py:176-180 with z1 = phi, y = theta, z2 = psi
        z1, y, z2 = np.radians([phi, theta, psi])
        cz1, sz1, cy, sy, cz2, sz2 = np.cos(z1), np.sin(z1), np.cos(y), np.sin(y), np.cos(z2), np.sin(z2)
        m = np.array([[cz1 * cy * cz2 - sz1 * sz2, cz1 * cy * sz2 + sz1 * cz2, -cz1 * sy],
                      [-sz1 * cy * cz2 - cz1 * sz2, -sz1 * cy * sz2 + cz1 * cz2, sz1 * sy],
                      [sy * cz2, sy * sz2, cy]])
        out_psi, out_theta, out_phi = geo.get_degrees_from_matrix(m)
        rows.append(np.concatenate([m.ravel(), [psi, theta, phi, out_psi, out_theta, out_phi]]))
    np.save(os.path.join(HERE, "euler_decode.npy"), np.array(rows))

    spa = []
    for _ in range(32):
        tilt, axis = rng.uniform(-60, 60), rng.uniform(80, 95)
        normal = rng.uniform(-30, 30, 3)
        ang = rng.uniform(0, 360, 3)
        # 3DAVG-style 4x4 refinement matrix: rotation (ZXZ) + translation
        rz1 = geo.vtk.rotation_matrix(np.radians(ang[0]), [0, 0, 1])
        rx = geo.vtk.rotation_matrix(np.radians(ang[1]), [1, 0, 0])
        rz2 = geo.vtk.rotation_matrix(np.radians(ang[2]), [0, 0, 1])
        mm = rz1 @ rx @ rz2
        mm[:3, 3] = rng.uniform(-4, 4, 3)
        res, pres = geo.spa_euler_angles(tilt, axis, normal, list(np.asarray(mm).ravel()), 0.0)
        spa.append(np.concatenate([[tilt, axis], normal, np.asarray(mm).ravel(), np.asarray(res, dtype=float), np.asarray(pres, dtype=float)]))
    np.save(os.path.join(HERE, "spa_euler.npy"), np.array(spa))

    csp = []
    for k in range(64):
        tilt, axis = rng.uniform(-70, 70), rng.uniform(-180, 180)
        ang = [rng.uniform(0, 360), rng.uniform(0, 180), rng.uniform(0, 360)]
        if k == 0:
            tilt, axis, ang = 0.0, 0.0, [30.0, 40.0, 50.0]      # SPA limit: no tilt geometry
        mm = np.eye(4)
        mm[:3, 3] = rng.uniform(-6, 6, 3)
        fp, nm = geo.csp_euler_angles(tilt, axis, [0.0, 0.0, 0.0], list(mm.ravel()), 0.0, ang)
        csp.append(np.concatenate([[tilt, axis], ang, mm[:3, 3], np.asarray(fp, dtype=float), np.asarray(nm, dtype=float)]))
    np.save(os.path.join(HERE, "csp_euler.npy"), np.array(csp))
    # ---- score shaping (scores.py:300-761), run in a scratch directory (it writes next to its input)
    import json
    import shutil
    import tempfile

    os.environ.setdefault("PYP_DIR", "/root/reference")
    from pyp.analysis import scores as S

    def shape_case(tag, tomo, cutoff, kw):
        n = 600
        d = np.zeros((n, 32))
        d[:, 0] = np.arange(1, n + 1)
        d[:, 1:4] = rng.uniform(0, 360, (n, 3))
        d[:, 6] = rng.uniform(5000, 40000, n)
        d[:, 7] = d[:, 6] - 200
        d[:, 10] = np.sort(rng.integers(0, 3, n))     # film id; rows are grouped by film as in pyp's merged tables
                                                      # (scores.py:348-361 builds the tilt column film by film)
        d[:, 11] = 100
        d[:, 14] = rng.normal(12, 4, n)
        d[:, 15] = 1.35
        nt = 5 if tomo else 1
        d[:, 26] = np.arange(n) // nt                 # PIND
        d[:, 27] = np.arange(n) % nt                  # TIND
        angles = [-40.0, -9.0, 0.0, 8.0, 41.0] if tomo else [0.0]
        tilts = {str(f): {str(t): a for t, a in enumerate(angles)} for f in range(3)}
        tmp = tempfile.mkdtemp()
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            par = csf.Parameters()
            par.set_data(data=d)
            par.to_binary("x_r01_02.cistem")
            json.dump(tilts, open("x_r01_02.json", "w"))
            S.shape_phase_residuals("x_r01_02.cistem", 1, 1, cutoff, kw["mindef"], kw["maxdef"], kw["firstframe"], kw["lastframe"],
                                    kw["mintilt"], kw["maxtilt"], kw["minazh"], kw["maxazh"], kw["minscore"], kw["maxscore"], 1.0,
                                    False, False, True, False, False, "x_r01_02_used.cistem")
            shutil.copy("x_r01_02.cistem", os.path.join(HERE, f"shape_{tag}_in.cistem"))
            shutil.copy("x_r01_02.json", os.path.join(HERE, f"shape_{tag}_in.json"))
            shutil.copy("x_r01_02_used.cistem", os.path.join(HERE, f"shape_{tag}_out.cistem"))
        finally:
            os.chdir(cwd)
            shutil.rmtree(tmp)
        with open(os.path.join(HERE, f"shape_{tag}_args.json"), "w") as f:
            json.dump(dict(kw, cutoff=cutoff), f)

    shape_case("spa", False, 0.8, dict(mindef=0.0, maxdef=30000.0, firstframe=0, lastframe=-1, mintilt=-90.0, maxtilt=90.0,
                                      minazh=10.0, maxazh=170.0, minscore=0.05, maxscore=0.98))
    shape_case("tomo", True, 0.7, dict(mindef=0.0, maxdef=100000.0, firstframe=0, lastframe=3, mintilt=-45.0, maxtilt=90.0,
                                      minazh=0.0, maxazh=180.0, minscore=0.0, maxscore=1.0))
    # ---- resolution schedule: postprocess.get_rhref (postprocess/core.py:16-55) on a synthetic FSC table
    from pyp import postprocess as PP

    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.makedirs(os.path.join(tmp, "maps"))
    os.makedirs(os.path.join(tmp, "scratch"))
    res = np.array([float(r) for r in 1.35 * 128 / np.arange(1, 60)])
    curves = [1.0 / (1.0 + (8.0 / res) ** 4), 1.0 / (1.0 + (6.0 / res) ** 4), 1.0 / (1.0 + (4.5 / res) ** 4)]
    np.savetxt(os.path.join(tmp, "maps", "ds_r01_fsc.txt"), np.column_stack([res] + curves))
    np.savetxt(os.path.join(tmp, "maps", "ds_r01_res.txt"), np.array([[2, 12.0], [3, 9.5], [4, 7.0]]))
    cases = []
    os.chdir(os.path.join(tmp, "scratch"))
    try:
        for rh, it in (("8:6:4", 2), ("8:6:4", 3), ("8:6:4", 7), ("0", 2), ("0", 3), ("0", 4), ("0", 5), (5.5, 3)):
            mp = {"refine_rhref": rh, "refine_dataset": "ds"}
            cases.append({"refine_rhref": rh, "iteration": it, "rhref": float(PP.get_rhref(mp, it))})
    finally:
        os.chdir(cwd)
    shutil.copy(os.path.join(tmp, "maps", "ds_r01_fsc.txt"), os.path.join(HERE, "rhref_fsc.txt"))
    shutil.copy(os.path.join(tmp, "maps", "ds_r01_res.txt"), os.path.join(HERE, "rhref_res.txt"))
    shutil.rmtree(tmp)
    with open(os.path.join(HERE, "rhref_cases.json"), "w") as f:
        json.dump(cases, f, indent=1)

    # ---- star tables: our writer read back by the reference's merge_star / read_star
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from pyp_b200.formats import cistem as our_cistem, star as our_star

    rows = our_cistem.read_parameters(os.path.join(HERE, "params_5x32.cistem"))
    our_star.write_star(os.path.join(HERE, "rows_a.star"), rows[:3])
    our_star.write_star(os.path.join(HERE, "rows_b.star"), rows[3:])
    merged = csf.merge_star([os.path.join(HERE, "rows_a.star"), os.path.join(HERE, "rows_b.star")])
    np.save(os.path.join(HERE, "rows_star_merged_by_reference.npy"), np.asarray(merged, dtype=np.float64))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
