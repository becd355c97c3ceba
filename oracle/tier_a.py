"""Tier-A probe and runner — TEST INFRASTRUCTURE ONLY (tests/, bench.py --impl reference).

SURVEY.md §0.6 / §8c: the reference's numerics live in closed binaries that ship as git-LFS stubs
(`external/cistem2/{refine3d,reconstruct3d,merge3d,local_merge3d}`, `external/CSP/csp`, ~133 bytes
each in /root/reference).  If real ELF files ever appear — a driver-written `baseline/_ref/`, or a pyp
installation at `$PYP_DIR/external/` — they ARE the oracle and the timed CPU baseline.  This module

  * finds them (`find_binaries`: ELF magic, size >> 133 B),
  * builds the stdin answer lists exactly as pyp does (`refine3d_answers`: frealign.py:3918-3994,
    `reconstruct3d_answers`: frealign.py:1780-1824, `merge3d_answers`: frealign.py:2075-2093; pinned
    against the heredocs produced by the reference's own builders in tests/golden/prompts_*.json),
  * and runs one refine3d -> reconstruct3d -> merge3d iteration the way pyp does (`run_iteration`):
    contiguous particle ranges of `increment = ceil(frames / cores)` rows (local_run.py:507-516), one
    concurrent single-thread process per range with the heredoc on stdin (`OMP_NUM_THREADS=NCPUS=1`,
    frealign.py:3183), range outputs merged like Parameters.merge, dumps summed by merge3d.

`CSPB_TIER_A_DIR=<dir>` points the probe at any directory of executables with those names without the
ELF check — used by the GPU test that drives this very runner against the drop-in front-ends in bin/.
"""
import glob
import math
import os
import subprocess
import time

import numpy as np

PROGRAMS = ("refine3d", "reconstruct3d", "merge3d", "local_merge3d", "csp")
_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)


def _is_elf(path, min_bytes=1 << 20):
    try:
        if os.path.getsize(path) < min_bytes:
            return False
        with open(path, "rb") as f:
            return f.read(4) == b"\x7fELF"
    except OSError:
        return False


def find_binaries():
    """{program: path} of real reference binaries, {} when only the LFS stubs (or nothing) exist."""
    found = {}
    forced = os.environ.get("CSPB_TIER_A_DIR")
    if forced:
        for prog in PROGRAMS:
            p = os.path.join(forced, prog)
            if os.path.isfile(p) and os.access(p, os.X_OK):
                found[prog] = p
        return found
    roots = [os.path.join(ROOT, "baseline", "_ref")]
    pyp_dir = os.environ.get("PYP_DIR")
    if pyp_dir:
        roots += [os.path.join(pyp_dir, "external", "cistem2"), os.path.join(pyp_dir, "external", "CSP")]
    for root in roots:
        if not os.path.isdir(root):
            continue
        for prog in PROGRAMS:
            if prog in found:
                continue
            for p in [os.path.join(root, prog)] + sorted(glob.glob(os.path.join(root, "**", prog), recursive=True)):
                if _is_elf(p):
                    found[prog] = p
                    break
    return found


def probe_report():
    """What the probe saw, for bench.py's cpu_baseline record and the logs."""
    b = find_binaries()
    have = all(k in b for k in ("refine3d", "reconstruct3d", "merge3d"))
    stubs = []
    pyp_dir = os.environ.get("PYP_DIR")
    if pyp_dir:
        for prog in PROGRAMS:
            for sub in ("cistem2", "CSP"):
                p = os.path.join(pyp_dir, "external", sub, prog)
                if os.path.isfile(p) and not _is_elf(p):
                    stubs.append(f"{p} ({os.path.getsize(p)} B, not ELF)")
    return {"tier": "A" if have else "B", "binaries": b, "stubs": stubs,
            "searched": ["baseline/_ref", "$PYP_DIR/external/{cistem2,CSP}" + ("" if pyp_dir else " (PYP_DIR unset)")]}


# ------------------------------------------------------------------------------ answer lists
def _yn(v):
    return "yes" if v else "no"


def refine3d_answers(stack, par, reference, statistics, name, first, last, pixel, mw, outer_radius, rlref, rhref, symmetry="C1",
                     stat="null", use_statistics=False, use_priors=False, signed_cc_limit="30.0", class_rhcls=None, search_radius=None,
                     search_rhref=None, dang="20.0", searchx="0", searchy="0", focus=("0", "0", "0", "0"), defocus_range=500,
                     iblow="1", global_search=False, local=True, mask=(1, 1, 1, 1, 1), matching=False, focus_mask=False,
                     refine_defocus=False, normalize=True, invert=False):
    """The 50 answers of frealign.py:3918-3994 (mrefine_version), in order.  `mask` = (psi, theta, phi, x, y)
    as the reference reads refine_mask; note its quirk: phi takes the flag of index 1 (frealign.py:3814-3817)."""
    ranger = "%07d_%07d" % (first, last)
    psi, theta, _, x, y = mask
    phi = mask[1]
    return [stack, par, stat, reference, statistics, _yn(use_statistics), _yn(use_priors),
            f"{name}_match.mrc_{ranger}", f"{name}_{ranger}.cistem", f"{name}_{ranger}_changes.cistem",
            symmetry, first, last, 1, pixel, mw, 0, outer_radius, rlref, rhref, signed_cc_limit,
            class_rhcls if class_rhcls is not None else rhref,
            search_radius if search_radius is not None else 1.5 * float(outer_radius),
            search_rhref if search_rhref is not None else rhref, dang, 20, searchx, searchy, *focus, defocus_range, "50.0", iblow,
            _yn(global_search), _yn(local), _yn(psi), _yn(theta), _yn(phi), _yn(x), _yn(y), _yn(matching), _yn(focus_mask),
            _yn(refine_defocus), _yn(normalize), _yn(invert), "no", "no", "no"]


def reconstruct3d_answers(stack, par, reference, name, first, last, pixel, mw, outer_radius, res_rec, dump1, dump2, symmetry="C1", stat="null",
                          bsc=2.0, score_weighting=False, dose=None, thresh=0, normalize=True, adjust=False, invert=False, crop=False,
                          per_particle=False, blurring=False, dump=True):
    """The answers of frealign.py:1780-1824 (split_reconstruction), in order; `dose` = None or the four
    extra answers (weights file, multiply yes/no, fraction, transition) of :1731-1753."""
    dose_lines = ["no"] if dose is None else ["yes", dose[0], _yn(dose[1]), dose[2], dose[3]]
    return [stack, par, stat, reference, f"{name}_map1.mrc", f"{name}_map2.mrc", "output.mrc", f"{name}_n{first}.res", symmetry, first, last,
            pixel, mw, 0, outer_radius, res_rec, 0, bsc, _yn(score_weighting), 0, -1, *dose_lines, thresh, 1, 1, _yn(normalize), _yn(adjust),
            _yn(invert), "no", _yn(crop), "yes", _yn(per_particle), "no", _yn(blurring), "no", _yn(dump), dump1, dump2, 1]


def merge3d_answers(out_stem, mw, outer_radius, seed1, seed2, count):
    """frealign.py:2075-2093 (merge_reconstructions)."""
    return [f"{out_stem}_half1.mrc", f"{out_stem}_half2.mrc", f"{out_stem}.mrc", f"{out_stem}_statistics.txt", mw, 0, outer_radius, seed1, seed2, count]


def heredoc(answers):
    return "\n".join(str(a) for a in answers) + "\n"


# ------------------------------------------------------------------------------ runner
def split_ranges(frames, cores):
    """local_run.py:507-516: 1-based inclusive ranges of `increment + 1` rows."""
    increment = math.ceil(frames / cores)
    return [(first, min(first + increment, frames)) for first in range(1, frames + 1, increment + 1)]


def _spawn(prog, answers, cwd, log):
    env = dict(os.environ, OMP_NUM_THREADS="1", NCPUS="1")
    with open(log, "ab") as lf:
        p = subprocess.Popen([prog], stdin=subprocess.PIPE, stdout=lf, stderr=subprocess.STDOUT, cwd=cwd, env=env)
    p.stdin.write(heredoc(answers).encode())
    p.stdin.close()
    return p


def _wait_all(procs, timeout):
    t_end = time.time() + timeout
    for p in procs:
        p.wait(timeout=max(1.0, t_end - time.time()))


def run_iteration(binaries, workdir, vol, stack, rows, pixel, symmetry="C1", mw=300.0, mask_radius=None, rlref=100.0, rhref=None,
                  cores=None, name="ds_r01", timeout=3600):
    """One pyp-style iteration with the given executables.  Returns dict(rows, half1, half2, map, statistics,
    refine_s, reconstruct_s, merge_s, ranges).  Inputs are written in the reference's file formats."""
    from pyp_b200.formats import cistem, mrc, statistics  # wire formats, pinned byte-for-byte against the reference's writers

    n = int(stack.shape[-1])
    cores = cores or os.cpu_count() or 1
    mask_radius = mask_radius if mask_radius is not None else 0.38 * n * pixel
    rhref = rhref if rhref is not None else 2.5 * pixel
    os.makedirs(os.path.join(workdir, "scratch"), exist_ok=True)
    mrc.write(os.path.join(workdir, "ds_stack.mrc"), np.ascontiguousarray(stack, dtype=np.float32), pixel)
    mrc.write(os.path.join(workdir, f"{name}.mrc"), np.ascontiguousarray(vol, dtype=np.float32), pixel)
    cistem.write_parameters(os.path.join(workdir, f"{name}.cistem"), rows)
    open(os.path.join(workdir, "statistics_r01.txt"), "w").close()
    ranges = split_ranges(int(rows.size), cores)
    # ---- refine3d: one process per range, all at once (mpi.py:44-48 joblib fan-out)
    t0 = time.perf_counter()
    procs = [_spawn(binaries["refine3d"],
                    refine3d_answers("ds_stack.mrc", f"{name}.cistem", f"{name}.mrc", "statistics_r01.txt", name, f, l, pixel, mw, mask_radius,
                                     rlref, rhref, symmetry), workdir, os.path.join(workdir, "refine3d.log")) for f, l in ranges]
    _wait_all(procs, timeout)
    refine_s = time.perf_counter() - t0
    outs = [os.path.join(workdir, f"{name}_{f:07d}_{l:07d}.cistem") for f, l in ranges]
    missing = [o for o in outs if not os.path.exists(o)]  # pyp's own success test: the files exist (frealign.py:3086-3094)
    if missing:
        raise RuntimeError(f"refine3d wrote {len(outs) - len(missing)} of {len(outs)} range files; log tail: "
                           + open(os.path.join(workdir, "refine3d.log"), errors="replace").read()[-800:])
    refined = cistem.merge(outs)
    cistem.write_parameters(os.path.join(workdir, f"{name}_used.cistem"), refined)
    # ---- reconstruct3d with dumps, one process per range
    t0 = time.perf_counter()
    procs = [_spawn(binaries["reconstruct3d"],
                    reconstruct3d_answers("ds_stack.mrc", f"{name}_used.cistem", f"{name}.mrc", name, f, l, pixel, mw, pixel * n / 2, 2 * pixel,
                                          f"scratch/{name}_map1_n{k}.mrc", f"scratch/{name}_map2_n{k}.mrc", symmetry),
                    workdir, os.path.join(workdir, "reconstruct3d.log")) for k, (f, l) in enumerate(ranges, start=1)]
    _wait_all(procs, timeout)
    reconstruct_s = time.perf_counter() - t0
    if "caught" in open(os.path.join(workdir, "reconstruct3d.log"), errors="replace").read():  # particle_cspt.py:812-818
        raise RuntimeError("reconstruct3d log contains 'caught'")
    # ---- merge3d
    t0 = time.perf_counter()
    p = _spawn(binaries["merge3d"], merge3d_answers(f"{name}_02", mw, pixel * n / 2, f"scratch/{name}_map1_n.mrc", f"scratch/{name}_map2_n.mrc", len(ranges)),
               workdir, os.path.join(workdir, "merge3d.log"))
    _wait_all([p], timeout)
    merge_s = time.perf_counter() - t0
    log = open(os.path.join(workdir, "merge3d.log"), errors="replace").read()
    if "Merge3D: Normal termination" not in log:  # frealign.py:2558
        raise RuntimeError("merge3d did not terminate normally: " + log[-800:])
    maps = [np.asarray(mrc.read(os.path.join(workdir, f"{name}_02{s}.mrc"))[1]) for s in ("_half1", "_half2", "")]
    st = statistics.read_statistics(os.path.join(workdir, f"{name}_02_statistics.txt"))
    return {"rows": refined, "half1": maps[0], "half2": maps[1], "map": maps[2], "statistics": st, "refine_s": refine_s,
            "reconstruct_s": reconstruct_s, "merge_s": merge_s, "ranges": ranges}
