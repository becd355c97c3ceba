"""Golden fixture for pyp_b200.csp_geometry.defocus_offset_from_center: the REFERENCE's
pyp.analysis.geometry.core.DefocusOffsetFromCenter (geometry/core.py:686-773) on random particles, tilts
and in-plane alignments.  Run in the build container only:  python tests/golden/make_golden_defocus.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401

import numpy as np  # noqa: E402

from pyp.analysis.geometry import core as geo  # noqa: E402


def main():
    rng = np.random.default_rng(9)
    rows = []
    for k in range(48):
        p = rng.uniform(0, 512, 3)
        c = np.array([256.0, 256.0, 128.0]) + rng.uniform(-4, 4, 3)
        tilt = rng.uniform(-65, 65)
        ax = np.radians(rng.uniform(-180, 180))
        T = [np.cos(ax), -np.sin(ax), np.sin(ax), np.cos(ax), rng.uniform(-20, 20), rng.uniform(-20, 20)]
        zo = rng.uniform(-30, 30)
        h = 1 if k % 2 else -1
        v = float(geo.DefocusOffsetFromCenter(list(p), list(c), tilt, T, zo, handedness=h))
        rows.append(np.concatenate([p, c, [tilt], T, [zo, h, v]]))
    np.save(os.path.join(HERE, "defocus_offset.npy"), np.array(rows))
    print("written defocus_offset.npy", len(rows))


if __name__ == "__main__":
    main()
