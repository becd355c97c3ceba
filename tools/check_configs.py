#!/usr/bin/env python
"""Run the other BASELINE.json configs at their real shapes (bounded particle counts) through the
public calls and report throughput: C1 128-px C1 refine+reconstruct, C4 384-px global search,
C5 512-px reconstruction with a 2x padded volume.  (C2 is bench.py, C3 is tools/bench_csp.py.)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyp_b200 import synth, synth_torch  # noqa: E402
from pyp_b200.engine import Engine  # noqa: E402
from pyp_b200.search_grid import search_grid  # noqa: E402


def setup(n, px, P, sym="C1", seed=0, radius_frac=0.35, sigma=2.0):
    dev = torch.device("cuda", 0)
    centres, amps, sigma = synth_torch.symmetric_phantom(n, sym, n_base=40 if sym == "C1" else 9, radius_frac=radius_frac, sigma=sigma)
    vol = synth_torch.volume(n, centres, amps, sigma, dev)
    rows = synth.make_rows(P, px, seed=seed + 1)
    stack = synth_torch.make_stack(n, centres, amps, sigma, rows, snr=0.05, seed=seed + 2, device=dev)
    return vol, rows, stack


def timed(eng, f, *a, **k):
    eng.sync()
    t0 = time.perf_counter()
    r = f(*a, **k)
    eng.sync()
    return time.perf_counter() - t0, r


def main():
    out = []
    eng = Engine(0)
    # ---- C1: 5k particles, 128-px box, C1, local refinement + reconstruction
    n, px, P = 128, 1.35, 5000
    vol, rows, stack = setup(n, px, P)
    cfg = Engine.refine_defaults(n, px)
    cfg.mask_radius = 0.38 * n * px
    eng.refine_configure(cfg)
    eng.set_symmetry("C1")
    eng.set_reference(vol)
    start = synth.perturb_rows(rows, 2.0, 1.0)
    host = stack.cpu().numpy()
    eng.recon_begin(Engine.recon_defaults(n, px))
    eng.refine_reconstruct(host, start)
    eng.recon_begin(Engine.recon_defaults(n, px))
    dt, (ref_rows, n_ev) = timed(eng, eng.refine_reconstruct, host, start)
    rec, h1, h2, stats = eng.recon_finalize(molecular_mass_kda=300.0)
    from tests_common import angular_distance
    out.append({"config": "C1 5k x 128 px, C1, refine3d + reconstruct3d (host stack in)", "seconds": dt, "scored_projections_per_s": n_ev / dt,
                "particles_per_s": P / dt, "median_angular_error_deg": float(np.median(angular_distance(ref_rows, rows))),
                "fsc_0.5_shell": int(np.argmax(stats[1:, 3] < 0.5) + 1)})
    del stack, host
    # ---- C4: 384-px box, global search (refine_mode 0, 20 degree grid)
    n, px, P = 384, 1.35, 512
    # a compact particle (radius 0.15 n = 78 A) searched to 30 A: the 20 degree grid of refine_dang's
    # default is then about twice the angular resolution of the search band, as in practice
    vol, rows, stack = setup(n, px, P, seed=10, radius_frac=0.15, sigma=4.0)
    cfg = Engine.refine_defaults(n, px)
    cfg.mask_radius = 0.25 * n * px
    cfg.global_search, cfg.local_refine = 1, 1
    cfg.search_high_res = 30.0
    cfg.search_range_x = cfg.search_range_y = 20.0
    cfg.best_matches = 20
    eng.refine_configure(cfg)
    eng.set_reference(vol)
    grid = search_grid(20.0, "C1")
    eng.set_search_grid(grid)
    eng.load_images(stack)
    start = rows.copy()
    for k in ("psi", "theta", "phi", "x_shift", "y_shift"):
        start[k] = 0
    dt, (got, _, n_ev) = timed(eng, eng.refine, start)
    out.append({"config": "C4 384 px global search, 20 deg grid (%d orientations), top-20 refined" % grid.shape[0], "particles": P, "seconds": dt,
                "scored_projections_per_s": n_ev / dt, "particles_per_s": P / dt,
                "median_angular_error_deg": float(np.median(angular_distance(got, rows)))})
    del stack
    # ---- C5: 512-px box, reconstruction into a 2x padded volume
    n, px, P = 512, 1.0, 1024
    vol, rows, stack = setup(n, px, P, seed=20)
    rc = Engine.recon_defaults(n, px)
    rc.pad = 2
    eng.recon_begin(rc)
    eng.recon_insert(stack[:64], torch.from_numpy(rows[:64].view(np.uint8).reshape(64, 128)).cuda())
    eng.recon_begin(rc)
    rows_dev = torch.from_numpy(rows.view(np.uint8).reshape(P, 128)).cuda()
    dt, _ = timed(eng, eng.recon_insert, stack, rows_dev)
    dtf, (rec, _, _, stats) = timed(eng, eng.recon_finalize, molecular_mass_kda=800.0, want_halves=False)
    v = vol.cpu().numpy()
    cc = float(np.corrcoef(rec.ravel(), v.ravel())[0, 1])
    out.append({"config": "C5 512 px reconstruct3d, pad 2 (1024^3 accumulators)", "particles": P, "insert_seconds": dt, "particles_per_s": P / dt,
                "finalize_seconds": dtf, "map_correlation_with_phantom": cc})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import common as tests_common  # noqa: E402
    sys.modules["tests_common"] = tests_common
    main()
