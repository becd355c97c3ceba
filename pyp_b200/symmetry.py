"""Point-group rotation matrices for the symmetry symbol pyp passes to the binaries
(prompt 11 of refine3d / prompt 9 of reconstruct3d, src/pyp/refine/frealign/frealign.py:3938,1799;
``particle_sym`` in config/pyp_config.toml).  Orientation conventions are the FREALIGN/cisTEM
ones: Cn/Dn principal axis along z, D's two-fold along x, T and O with two-/four-folds along
x, y, z, I with two-folds along x, y, z (I2).  *(external knowledge — the reference tree only
carries the symbol)*."""
import math

import numpy as np


def _rot(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    x, y, z = axis
    c, s = math.cos(angle), math.sin(angle)
    C = 1.0 - c
    return np.array(
        [
            [c + x * x * C, x * y * C - z * s, x * z * C + y * s],
            [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
            [z * x * C - y * s, z * y * C + x * s, c + z * z * C],
        ]
    )


def _closure(gens, limit=120):
    mats = [np.eye(3)]
    frontier = [np.eye(3)]
    while frontier:
        new = []
        for m in frontier:
            for g in gens:
                c = g @ m
                if not any(np.allclose(c, k, atol=1e-9) for k in mats):
                    mats.append(c)
                    new.append(c)
                    if len(mats) > limit:
                        raise ValueError("symmetry group did not close")
        frontier = new
    return mats


def symmetry_matrices(symbol: str) -> np.ndarray:
    """Return (n_ops, 3, 3) float32 rotation matrices for ``symbol`` (C1, C7, D2, T, O, I ...)."""
    s = symbol.strip().upper()
    if not s:
        raise ValueError("empty symmetry symbol")
    kind, order = s[0], s[1:]
    if kind == "C":
        n = int(order or 1)
        if n < 1:
            raise ValueError(symbol)
        mats = [_rot([0, 0, 1], 2 * math.pi * k / n) for k in range(n)]
    elif kind == "D":
        n = int(order)
        if n < 1:
            raise ValueError(symbol)
        cn = [_rot([0, 0, 1], 2 * math.pi * k / n) for k in range(n)]
        c2 = _rot([1, 0, 0], math.pi)
        mats = cn + [c2 @ m for m in cn]
    elif kind == "T" and order == "":
        mats = _closure([_rot([0, 0, 1], math.pi), _rot([1, 1, 1], 2 * math.pi / 3)])
        assert len(mats) == 12
    elif kind == "O" and order == "":
        mats = _closure([_rot([0, 0, 1], math.pi / 2), _rot([1, 1, 1], 2 * math.pi / 3)])
        assert len(mats) == 24
    elif kind == "I" and order in ("", "2"):
        phi = (1 + math.sqrt(5)) / 2
        mats = _closure([_rot([0, 0, 1], math.pi), _rot([1, 1, 1], 2 * math.pi / 3), _rot([0, 1, phi], 2 * math.pi / 5)])
        assert len(mats) == 60
    else:
        raise ValueError(f"unsupported symmetry symbol {symbol!r}")
    return np.ascontiguousarray(np.stack(mats).astype(np.float32))
