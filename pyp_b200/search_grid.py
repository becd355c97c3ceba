"""Orientation grid of refine3d's global search (prompt 25 angular step `refine_dang`,
src/pyp/refine/frealign/frealign.py:3957).  Near-uniform coverage of the asymmetric unit in
(phi, theta) times an in-plane psi sweep, the scheme of cisTEM's EulerSearch *(external
knowledge; limits per point group are conservative supersets of the asymmetric unit)*."""
import math

import numpy as np


def _limits(symbol):
    s = symbol.strip().upper()
    kind, order = s[0], s[1:]
    if kind == "C":
        return 360.0 / max(1, int(order or 1)), 180.0
    if kind == "D":
        return 360.0 / max(1, int(order)), 90.0
    if kind == "T":
        return 180.0, 90.0
    if kind == "O":
        return 90.0, 90.0
    if kind == "I":
        return 72.0, 90.0
    raise ValueError(f"unsupported symmetry symbol {symbol!r}")


def search_grid(angular_step, symbol="C1", psi_step=None):
    """(n, 3) float32 array of (psi, theta, phi) in degrees."""
    step = float(angular_step)
    if step <= 0:
        raise ValueError("angular step must be positive")
    phi_max, theta_max = _limits(symbol)
    n_theta = max(1, int(round(theta_max / step)))
    psi_step = float(psi_step or step)
    n_psi = max(1, int(round(360.0 / psi_step)))
    views = []
    for k in range(n_theta + 1):
        theta = k * theta_max / n_theta
        st = math.sin(math.radians(theta))
        n_phi = 1 if st < 1e-6 else max(1, int(math.ceil(phi_max * st / step)))
        for m in range(n_phi):
            views.append((theta, m * phi_max / n_phi))
    out = np.zeros((len(views) * n_psi, 3), dtype=np.float32)
    r = 0
    for theta, phi in views:
        for q in range(n_psi):
            out[r] = (q * 360.0 / n_psi, theta, phi)
            r += 1
    return out
