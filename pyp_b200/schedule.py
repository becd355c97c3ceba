"""Per-iteration parameter schedule of the refinement: colon lists and the
high-resolution limit handed to refine3d / refine_ctf / csp.

Restates src/pyp/system/project_params.py:362-373 (`param`) and src/pyp/postprocess/core.py:16-55
(`get_rhref`): a positive `refine_rhref` is used as is; a negative one is jittered by +-4 % in
reciprocal space; zero derives the limit from the FSC curve of the previous iterations
(`<maps>/<dataset>_r01_fsc.txt`, column `iteration-2`) at the previous limit
(`<maps>/<dataset>_r01_res.txt`, row `iteration-3`, else 16 A): if the FSC there is above the cutoff the
limit advances by int(FSC/cutoff) shells, otherwise it stays.  Pinned by tests/golden/rhref_cases.json.
"""
import os
import random

import numpy as np


def param(value, iteration):
    if isinstance(value, str):
        listed = value.split(":")
        return listed[min(iteration - 2, len(listed) - 1)]
    return value


def get_rhref(parameters, iteration, maps_dir="../maps", cutoff=0.143, rng=random):
    rhref = float(param(parameters["refine_rhref"], iteration))
    if rhref > 0:
        return rhref
    if rhref < 0:
        spread = 1.0 / rhref / 25.0
        return -1.0 / (1.0 / rhref + rng.uniform(-spread, spread))
    fsc_file = os.path.join(maps_dir, "%s_r01_fsc.txt" % parameters["refine_dataset"])
    if not (os.path.isfile(fsc_file) and iteration > 2):
        return 16
    fsc = np.loadtxt(fsc_file, ndmin=2, dtype=float)
    res_file = os.path.join(maps_dir, "%s_r01_res.txt" % parameters["refine_dataset"])
    prev = np.loadtxt(res_file, ndmin=2, dtype=float)[iteration - 3, 1] if os.path.exists(res_file) else 16.0
    x, y = fsc[:, 0], fsc[:, iteration - 2]
    order = np.argsort(x)
    if not (x.min() <= prev <= x.max()):
        raise ValueError("A value in x_new is outside the interpolation range.")  # scipy.interpolate.interp1d's behaviour
    current = float(np.interp(prev, x[order], y[order]))
    if current > cutoff:
        shells = int(current / cutoff)
        current_shell = int(np.argmin(np.abs(fsc[:, 0] - prev)))
        return fsc[current_shell + shells, 0]
    return prev
