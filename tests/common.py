"""Shared builders for the parity tests (same seeded inputs for the CUDA path and the oracle)."""
import numpy as np

from pyp_b200 import synth
from pyp_b200.engine import Engine, ROW_DTYPE


def small_case(n=64, n_part=24, px=1.35, snr=0.1, n_blobs=60, seed=1, shift_px=3.0):
    ph = synth.Phantom(n, n_blobs=n_blobs, sigma=1.5)
    vol = ph.volume()
    rows = synth.make_rows(n_part, px, seed=seed, shift_px=shift_px)
    stack = synth.make_stack(ph, rows, snr=snr, seed=seed + 1)
    return ph, vol, rows, stack


def refine_cfg(n, px, high_res=None, **kw):
    cfg = Engine.refine_defaults(n, px)
    cfg.mask_radius = 0.38 * n * px
    cfg.low_res_limit = 60.0
    cfg.high_res_limit = high_res if high_res else 4.0 * px
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def pose_of(r):
    return [r["psi"], r["theta"], r["phi"], r["x_shift"], r["y_shift"], 0.0]


def angular_distance(a, b):
    out = []
    for x, y in zip(a, b):
        p = synth.euler_matrix(x["psi"], x["theta"], x["phi"])
        q = synth.euler_matrix(y["psi"], y["theta"], y["phi"])
        out.append(np.degrees(np.arccos(np.clip((np.trace(p.T @ q) - 1) / 2, -1, 1))))
    return np.array(out)
