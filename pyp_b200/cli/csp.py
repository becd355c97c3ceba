"""csp drop-in: same argv as external/CSP/csp (src/pyp/system/local_run.py:364-376,392-404,
451-463), numerics on the GPU.

    csp <par.cistem> <par_extended.cistem> <mode> <first> <last> <flag> <images> <stack>  > log

cwd is the film's scratch directory; numeric options come from ``./.pyp_config.toml`` which the
driver re-saves right before every mode (src/pyp/align/core.py:1055); the reference map is
``$PYP_SCRATCH/<data_set>_frames_CSP_01.mrc`` (align/core.py:921-931).

Modes (align/core.py:1015-1023 after the mapping at local_run.py:332-335,411-431):
  -2  extract particles first..last from the tilt series `images` into `stack`
   3 with flag 0 and a frame list as `images`: frame (movie) refinement of the projections of particles first..last
      (run_frame_shifts); every other mode with a frame list runs as below on the stack — the patch-based region commands
      of local_run.py:337-404 hand region parameter files with mapped modes, refine_frames as flag and last = -1 for all
      micrographs of the patch
   0/3/6  tilt angle+axis / tilt shifts / both, for TIND first..last (last = -1: all)
   1/2/5  particle angles / shifts / both, for PIND first..last
   4  per-tilt defocus offset
Outputs: ``<par minus .cistem>_<first:06d>_<last:06d>.cistem`` with the rows of the refined
entities and ``..._extended.cistem`` with the refined entities only, which is what
merge_alignment_parameters globs and Parameters.merge overlays on the input tables
(src/pyp/refine/csp/particle_cspt.py:95-138; cistem_star_file.py:656-692).
"""
import os
import sys
import time

import numpy as np

from ..formats import cistem, mrc
from .prompts import PromptError, banner, pick_device

try:
    import tomllib
except ImportError:  # pragma: no cover
    tomllib = None

# defaults of config/pyp_config.toml for the keys the binary needs
DEFAULTS = {
    "csp_UseImagesForRefinementMin": 0, "csp_UseImagesForRefinementMax": 20, "csp_RefineProjectionCutoff": 0,
    "csp_NumberOfRandomIterations": 0, "csp_OptimizerMaxIter": 5, "csp_GridSearch": False, "csp_AngleStep": 20.0, "csp_ShiftStep": 6.0,
    "csp_ToleranceMicrographTiltAngles": 1.5, "csp_ToleranceMicrographTiltAxisAngles": 1.0, "csp_ToleranceMicrographShifts": 100.0,
    "csp_ToleranceParticlesPhi": 30.0, "csp_ToleranceParticlesPsi": 30.0, "csp_ToleranceParticlesTheta": 30.0,
    "csp_ToleranceParticlesShifts": 20.0, "csp_ToleranceMicrographDefocus1": 750.0,
    "refine_rlref": 100.0, "refine_rhref": "10", "refine_iter": 2, "refine_fboost": False, "refine_fboostlim": 0.0,
    "refine_iblow": "1", "extract_bin": 1, "data_bin": 1, "particle_sym": "C1",
}


from ..schedule import get_rhref, param  # noqa: E402  (per-iteration colon lists, project_params.py:362-373)


def load_config(path=".pyp_config.toml"):
    cfg = dict(DEFAULTS)
    if os.path.exists(path):
        if tomllib is None:
            raise ValueError("tomllib is unavailable; cannot read .pyp_config.toml")
        with open(path, "rb") as f:
            cfg.update(tomllib.load(f))
    return cfg


def csp_cfg_from(config, mode):
    from ..engine import Engine

    c = Engine.csp_defaults(mode)
    it = int(config.get("refine_iter", 2))
    c.window_min = int(param(config["csp_UseImagesForRefinementMin"], it))
    c.window_max = int(param(config["csp_UseImagesForRefinementMax"], it))
    c.iterations = int(config["csp_OptimizerMaxIter"])
    c.random_evals = int(config["csp_NumberOfRandomIterations"])
    # bin/csp_GS is the grid-search build of the binary (align/core.py:696-701)
    c.grid_search = 1 if (config["csp_GridSearch"] or os.environ.get("CSPB_CSP_GRID")) else 0
    c.angle_step, c.shift_step = float(config["csp_AngleStep"]), float(config["csp_ShiftStep"])
    c.tol_particle_psi = float(config["csp_ToleranceParticlesPsi"])
    c.tol_particle_theta = float(config["csp_ToleranceParticlesTheta"])
    c.tol_particle_phi = float(config["csp_ToleranceParticlesPhi"])
    c.tol_particle_shift = float(config["csp_ToleranceParticlesShifts"])
    c.tol_tilt_angle = float(config["csp_ToleranceMicrographTiltAngles"])
    c.tol_tilt_axis = float(config["csp_ToleranceMicrographTiltAxisAngles"])
    c.tol_tilt_shift = float(config["csp_ToleranceMicrographShifts"])
    c.tol_defocus = float(config["csp_ToleranceMicrographDefocus1"])
    c.min_projections = int(config["csp_RefineProjectionCutoff"])
    c.seed = int(config.get("csp_seed", 0)) & 0xFFFFFFFF
    return c


def refine_cfg_from(config, box, pixel):
    from ..engine import Engine

    it = int(config.get("refine_iter", 2))
    cfg = Engine.refine_defaults(box, pixel)
    cfg.low_res_limit = float(param(config["refine_rlref"], it))
    sched = {"refine_rhref": config["refine_rhref"], "refine_dataset": config.get("refine_dataset", config.get("data_set", ""))}
    cfg.high_res_limit = max(float(get_rhref(sched, it, maps_dir=os.path.join("frealign", "maps"))), 2.0 * pixel)  # postprocess/core.py:16-55
    rad = config.get("particle_rad")
    if rad:
        cfg.mask_radius = float(rad)
    cfg.signed_cc_limit = float(config["refine_fboostlim"]) if config.get("refine_fboost") else 30.0  # frealign.py:3879-3881
    cfg.pad = 2 if float(param(config["refine_iblow"], it)) >= 1.5 else 1
    return cfg


def out_paths(par, first, last):
    """particle_cspt.py:113-122: `<par>.strip('.cistem') + '_%06d_%06d' + '.cistem'`."""
    stem = par[:-len(".cistem")] if par.endswith(".cistem") else par
    tag = f"{stem}_{int(first):06d}_{int(last):06d}"
    return tag + ".cistem", tag + "_extended.cistem"


def reference_path(config):
    scratch = os.environ.get("PYP_SCRATCH", ".")
    return os.path.join(scratch, f"{config.get('data_set', 'dataset')}_frames_CSP_01.mrc")


def entity_rows(rows, particles, tilts, mode, first, last):
    """Index of the rows that belong to the entities first..last (PIND or TIND; last < 0 = open)."""
    key = rows["pind"] if mode in (1, 2, 5, -2) else rows["tind"]
    sel = key >= first
    if last >= 0:
        sel &= key <= last
    return np.nonzero(sel)[0]


def run_extract(par, mode, first, last, images, stack, config, out, session=None):
    from .session import Session

    session = session or Session()

    rows_all = cistem.read_parameters(par)
    idx = entity_rows(rows_all, None, None, -2, first, last)
    rows = rows_all[idx]
    rows = rows[np.argsort(rows["position_in_stack"], kind="stable")]
    box = int(config.get("extract_box", 0))
    binning = int(config.get("extract_bin", 1))
    if box <= 0:
        raise ValueError("extract_box missing from .pyp_config.toml")
    if rows.size == 0:
        out.write(f"csp: no projections for particles {first}..{last}; nothing extracted\n")
        return
    _, series = mrc.read(images)
    eng = session.engine(first + 1, max(1, last - first + 1))
    got = eng.csp_extract(np.ascontiguousarray(series, dtype=np.float32), rows, box * binning, binning)
    session.release()
    os.makedirs(os.path.dirname(stack) or ".", exist_ok=True)
    mrc.write(stack, got, pixel_size=float(rows["pixel_size"][0]))
    out.write(f"Extracted {rows.size} projections of particles {first}..{last} into {stack} ({box} px, bin {binning})\n")


def run_refine(par, ext, mode, first, last, stack, config, out, session=None):
    from .session import Session

    session = session or Session()

    t0 = time.time()
    rows_all = cistem.read_parameters(par)
    particles, tilts = cistem.read_extended(ext)
    idx = entity_rows(rows_all, particles, tilts, mode, first, last)
    out_par, out_ext = out_paths(par, first, last)
    if idx.size == 0:
        # no file for an empty range: the reference's reader raises 'Binary file is broken.' on a zero-row payload
        # and its writer asserts rows > 0 (cistem_star_file.py:694-776), and merge_alignment_parameters loads every
        # `*_??????_??????.cistem` it globs (particle_cspt.py:113-138) — an empty file would crash pyp's merge
        out.write(f"csp: nothing to refine for {first}..{last}; no output written\n")
        return
    rows = rows_all[idx]
    order = np.argsort(rows["position_in_stack"], kind="stable")
    rows = rows[order]
    hdr = mrc.read_header(stack)
    box = hdr["nx"]
    pixel = float(rows["pixel_size"][0])
    ref_path = reference_path(config)
    eng = session.engine(first + 1, max(1, last - first + 1))
    eng.ensure_reference(refine_cfg_from(config, box, pixel), ref_path, lambda: mrc.read(ref_path)[1])
    pos = rows["position_in_stack"].astype(np.int64)
    if pos.min() < 1 or pos.max() > hdr["nz"]:
        raise ValueError(f"POSITION_IN_STACK {pos.min()}..{pos.max()} outside the stack (1..{hdr['nz']})")
    chunk = 16384
    for s in range(0, rows.size, chunk):
        eng.load_images(session.images(stack, pos[s:s + chunk]), append=s > 0)
    ccfg = csp_cfg_from(config, mode)
    new_rows, new_p, new_t, n_evals = eng.csp_run(rows, particles, tilts, ccfg, first, last)
    session.release()
    cistem.write_parameters(out_par, new_rows)
    if mode in (1, 2, 5):
        sel = (particles["pind"] >= first) & ((particles["pind"] <= last) if last >= 0 else True)
        cistem.write_extended(out_ext, new_p[sel], new_t[:0])
    else:
        sel = (tilts["tind"] >= first) & ((tilts["tind"] <= last) if last >= 0 else True)
        cistem.write_extended(out_ext, new_p[:0], new_t[sel])
    dt = time.time() - t0
    out.write(f"mode {mode}: entities {first}..{last}, {rows.size} projections, exposures {ccfg.window_min}..{ccfg.window_max}\n")
    out.write(f"mean score {float(rows['score'].mean()):.4f} -> {float(new_rows['score'].mean()):.4f}, "
              f"{n_evals} projections scored in {dt:.2f} s\n")


def run_frame_shifts(par, ext, first, last, stack, config, out, session=None):
    """Frame (movie) refinement as pyp drives it without patches: `csp <par> <ext> 3 <p> <p> 0 frames_csp.txt <stack>`, one
    particle per process (local_run.py:434-439).  The stack then holds one projection per (particle, tilt, frame) and the
    rows carry FIND and FSHIFT_X/Y.  Ours (SEMANTICS.md §11, the closed binary's model is not public): for the rows of
    particles first..last every projection's in-plane shift is refined on the scorer with the angles held (the local
    optimiser of §7c restricted to x, y) within +-csp_ToleranceMicrographShifts of its input; the shift found goes to
    X_SHIFT / Y_SHIFT and its change is added to FSHIFT_X / FSHIFT_Y.  Tables of the extended file are untouched."""
    from .session import Session

    session = session or Session()
    t0 = time.time()
    rows_all = cistem.read_parameters(par)
    idx = entity_rows(rows_all, None, None, 5, first, last)
    out_par, out_ext = out_paths(par, first, last)
    if idx.size == 0:
        out.write(f"csp: no projections of particles {first}..{last}; no output written\n")
        return
    rows = rows_all[idx]
    rows = rows[np.argsort(rows["position_in_stack"], kind="stable")]
    hdr = mrc.read_header(stack)
    box, pixel = hdr["nx"], float(rows["pixel_size"][0])
    ref_path = reference_path(config)
    eng = session.engine(first + 1, max(1, last - first + 1))
    cfg = refine_cfg_from(config, box, pixel)
    cfg.refine_psi = cfg.refine_theta = cfg.refine_phi = 0
    cfg.refine_x = cfg.refine_y = 1
    eng.ensure_reference(cfg, ref_path, lambda: mrc.read(ref_path)[1])
    pos = rows["position_in_stack"].astype(np.int64)
    if pos.min() < 1 or pos.max() > hdr["nz"]:
        raise ValueError(f"POSITION_IN_STACK {pos.min()}..{pos.max()} outside the stack (1..{hdr['nz']})")
    for s in range(0, rows.size, 16384):
        eng.load_images(session.images(stack, pos[s:s + 16384]), append=s > 0)
    refined, _, n_evals = eng.refine(rows)
    session.release()
    tol = float(config["csp_ToleranceMicrographShifts"])
    new = refined.copy()
    for k, f in (("x_shift", "fshift_x"), ("y_shift", "fshift_y")):
        d = np.clip(refined[k] - rows[k], -tol, tol)
        new[k] = rows[k] + d
        new[f] = rows[f] + d
    cistem.write_parameters(out_par, new)
    p_tab, t_tab = cistem.read_extended(ext)
    cistem.write_extended(out_ext, p_tab[:0], t_tab[:0])
    out.write(f"frame shifts: particles {first}..{last}, {rows.size} projections (frames), tolerance {tol:g} A\n")
    out.write(f"mean score {float(rows['score'].mean()):.4f} -> {float(new['score'].mean()):.4f}, {n_evals} projections scored in {time.time() - t0:.2f} s\n")


def parse_argv(argv):
    """The eight arguments pyp hands over (src/pyp/system/local_run.py:306-467 create_csp_split_commands):
    parameter file, extended file, mode, first, last, flag (extract_frame / refine_frames), images, stack."""
    if len(argv) != 8:
        raise PromptError("usage: csp <par.cistem> <par_extended.cistem> <mode> <first> <last> <flag> <images> <stack>")
    par, ext, mode, first, last, flag, images, stack = argv
    return {"par": par, "ext": ext, "mode": int(float(mode)), "first": int(first), "last": int(last), "flag": int(float(flag)),
            "images": images, "stack": stack}


def main(argv=None, out=sys.stdout, session=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    try:
        a = parse_argv(argv)
        par, ext, mode, first, last, images, stack = a["par"], a["ext"], a["mode"], a["first"], a["last"], a["images"], a["stack"]
        out.write(banner("CSP"))
        frames = images.endswith(".txt")  # `images` = frames_csp.txt: frame (movie) refinement, local_run.py:320-323
        config = load_config()
        if mode == -1:
            pass
        elif frames and mode == -2:
            # cutting boxes out of movie frames needs the frame files behind the list, whose layout only the closed
            # binary knows: fail loudly instead of treating the list as a tilt series
            raise PromptError("csp: extraction (mode -2) from a frame list is not implemented in cspb200")
        elif frames and mode == 3 and a["flag"] == 0:
            # global frame refinement, one particle per process (local_run.py:434-439)
            run_frame_shifts(par, ext, first, last, stack, config, out, session)
        elif mode == -2:
            run_extract(par, mode, first, last, images, stack, config, out, session)
        elif mode in (0, 1, 2, 3, 4, 5, 6):
            run_refine(par, ext, mode, first, last, stack, config, out, session)
        else:
            raise PromptError(f"csp: unknown mode {mode}")
        out.write("\nCSP: Normal termination\n")
    except (PromptError, ValueError, OSError, RuntimeError, ImportError) as e:
        sys.stderr.write(f"csp: caught error: {e}\n")
        out.write("PYP (cspswarm) failed\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
