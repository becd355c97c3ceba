"""Golden fixture for the particle normalisation: the REFERENCE's pyp.analysis.image.normalize_image
(src/pyp/analysis/image.py:320-338,406-417) on a random image for radii inside the box, touching it and
beyond the half box (clamped there).  Run in the build container only (imports /root/reference):

    python tests/golden/make_golden_normalize.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401

import numpy as np  # noqa: E402

os.environ.setdefault("PYP_DIR", "/root/reference")
from pyp.analysis import image as IM  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    n, px = 48, 1.35
    img = (rng.normal(3.0, 2.0, (n, n)) + 0.05 * np.arange(n)[None, :]).astype(np.float32)
    radii_a = [10.0 * px, 20.0 * px, 24.0 * px, 30.0 * px, 40.0 * px]   # the last two exceed the half box
    out = np.stack([IM.normalize_image(img.copy(), r, px, 1) for r in radii_a]).astype(np.float32)
    np.savez(os.path.join(HERE, "normalize_image.npz"), image=img, pixel=px, radii_angstrom=np.array(radii_a), normalized=out)
    print("written normalize_image.npz", out.shape)


if __name__ == "__main__":
    main()
