#!/usr/bin/env python
"""CSP throughput on a synthetic tilt series shaped like BASELINE configs[2] (sub-tomogram CSP:
particles x 41 tilts, 128-px box, per-tilt defocus): particle mode 5 and micrograph mode 6 through
the public call (host tables in, host tables out), scored projections / s."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyp_b200 import synth, synth_torch  # noqa: E402
from pyp_b200.engine import Engine  # noqa: E402


def main():
    n_part = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    n, px = 128, 1.35
    dev = torch.device("cuda", 0)
    centres, amps, sigma = synth_torch.symmetric_phantom(n, "C1", n_base=60)
    vol = synth_torch.volume(n, centres, amps, sigma, dev)
    rows, particles, tilts = synth.make_tilt_series(n_part, px, seed=1)
    stack = synth_torch.make_stack(n, centres, amps, sigma, rows, snr=0.05, seed=3, device=dev)
    eng = Engine(0)
    cfg = Engine.refine_defaults(n, px)
    cfg.low_res_limit, cfg.high_res_limit, cfg.mask_radius = 100.0, 2.5 * px, 0.38 * n * px
    eng.refine_configure(cfg)
    eng.set_reference(vol)
    eng.load_images(stack)
    start_p = synth.perturb_particles(particles, 2.0, 1.5)
    t0 = time.perf_counter()
    start_rows = synth.rows_from_tables(rows, particles, tilts, start_p, tilts) if n_part <= 500 else rows
    out = []
    for mode, label in ((5, "particles (mode 5)"), (6, "micrographs (mode 6)")):
        ccfg = Engine.csp_defaults(mode)
        ccfg.window_max = 20
        ccfg.iterations = 5
        eng.csp_run(start_rows, start_p if n_part <= 500 else particles, tilts, ccfg)  # warm-up
        eng.sync()
        times = []
        for _ in range(7):
            t0 = time.perf_counter()
            r, p, t, n_ev = eng.csp_run(start_rows, start_p if n_part <= 500 else particles, tilts, ccfg)
            times.append(time.perf_counter() - t0)
        dt = float(np.median(times))
        out.append({"mode": label, "projections": int(rows.size), "evals": n_ev, "seconds": dt, "scored_projections_per_s": n_ev / dt,
                    "n_band": eng.band_counts()[0]})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
