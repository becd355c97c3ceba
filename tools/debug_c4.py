#!/usr/bin/env python
"""Small C4-shaped global search (debug helper): python tools/debug_c4.py <box> <particles> <K> [with_c1_first]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from check_configs import setup  # noqa: E402
from pyp_b200 import synth  # noqa: E402
from pyp_b200.engine import Engine  # noqa: E402
from pyp_b200.search_grid import search_grid  # noqa: E402

n, P, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
px = 1.35
eng = Engine(0)
if len(sys.argv) > 4:
    vol, rows, stack = setup(128, px, 256)
    cfg = Engine.refine_defaults(128, px)
    cfg.mask_radius = 0.38 * 128 * px
    eng.refine_configure(cfg)
    eng.set_symmetry("C1")
    eng.set_reference(vol)
    eng.recon_begin(Engine.recon_defaults(128, px))
    eng.refine_reconstruct(stack.cpu().numpy(), synth.perturb_rows(rows, 2.0, 1.0))
    print("c1 part ok", flush=True)
vol, rows, stack = setup(n, px, P, seed=10, radius_frac=0.15, sigma=4.0)
cfg = Engine.refine_defaults(n, px)
cfg.mask_radius = 0.25 * n * px
cfg.global_search, cfg.local_refine = 1, 1
cfg.search_high_res = 30.0
cfg.search_range_x = cfg.search_range_y = 20.0
cfg.best_matches = K
eng.refine_configure(cfg)
eng.set_reference(vol)
eng.set_search_grid(search_grid(20.0, "C1"))
eng.load_images(stack)
start = rows.copy()
for k in ("psi", "theta", "phi", "x_shift", "y_shift"):
    start[k] = 0
got, _, n_ev = eng.refine(start)
eng.sync()
print("ok", n, P, K, n_ev, float(got["score"].mean()), flush=True)
