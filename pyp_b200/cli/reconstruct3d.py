"""reconstruct3d drop-in: stdin contract of external/cistem2/reconstruct3d as pyp drives it
(src/pyp/refine/frealign/frealign.py:1780-1824), including the
conditional 4-line expansion of the dose-weighting answer (:1731-1753)."""
import sys
import time

import numpy as np

from ..formats import cistem, dump, mrc
from .prompts import Answers, PromptError, banner, pick_device, write_notes
from .refine3d import select_rows


def parse(ans: Answers):
    p = {}
    p["stack"] = ans.text("input particle images")
    p["parameters"] = ans.text("input cisTEM parameter file")
    p["global_stat"] = ans.text("global statistics parameter file")
    p["reference"] = ans.text("input reconstruction")
    p["out_map1"] = ans.text("output reconstruction 1")
    p["out_map2"] = ans.text("output reconstruction 2")
    p["out_filtered"] = ans.text("output filtered reconstruction")
    p["out_statistics"] = ans.text("output resolution statistics")
    p["symmetry"] = ans.text("particle symmetry")
    p["first"] = ans.integer("first particle to include")
    p["last"] = ans.integer("last particle to include")
    p["pixel_size"] = ans.number("pixel size of images")
    p["molecular_mass"] = ans.number("molecular mass of particle (kDa)")
    p["inner_mask_radius"] = ans.number("inner mask radius")
    p["outer_mask_radius"] = ans.number("outer mask radius")
    p["resolution_limit"] = ans.number("resolution limit for reconstruction")
    p["resolution_ref"] = ans.number("resolution limit of reference")
    p["score_bfactor"] = ans.number("particle weighting factor (score to B-factor)")
    p["score_weighting"] = ans.yesno("score weighting")
    p["min_tilt_score"] = ans.number("minimum tilt-particle score")
    p["max_tilt_score"] = ans.number("maximum tilt-particle score")
    p["dose_weighting"] = ans.yesno("dose weighting")
    if p["dose_weighting"]:
        p["dose_weights_file"] = ans.text("external weights file")
        p["dose_multiply"] = ans.yesno("multiply weights")
        p["dose_fraction"] = ans.number("dose weighting fraction")
        p["dose_transition"] = ans.number("dose weighting transition")
    p["score_threshold"] = ans.number("score threshold")
    p["smoothing"] = ans.number("tuning parameter: smoothing factor")
    p["padding"] = ans.number("tuning parameter: padding factor")
    p["normalize"] = ans.yesno("normalize particles")
    p["adjust_scores"] = ans.yesno("adjust scores for defocus dependence")
    p["invert"] = ans.yesno("invert particle contrast")
    p["exclude_edges"] = ans.yesno("exclude images with blank edges")
    p["crop"] = ans.yesno("crop particle images")
    p["split_even_odd"] = ans.yesno("FSC calculation with even/odd particles")
    p["per_particle_split"] = ans.yesno("per-particle splitting")
    p["center_mass"] = ans.yesno("center mass")
    p["likelihood_blurring"] = ans.yesno("apply likelihood blurring")
    p["threshold_input"] = ans.yesno("threshold input reconstruction")
    p["dump"] = ans.yesno("dump intermediate arrays")
    p["dump1"] = ans.text("output dump filename for odd particles")
    p["dump2"] = ans.text("output dump filename for even particles")
    p["max_threads"] = ans.integer("max threads") if not ans.done() else 1
    return p


def ignored_answers(p):
    """Answers of the list that this implementation reads and does not act on (each gets a log line)."""
    return [
        (p["inner_mask_radius"] != 0, f"inner mask radius {p['inner_mask_radius']:g} A accepted and ignored (pyp always passes 0)"),
        (p["resolution_ref"] != 0, f"resolution limit of reference {p['resolution_ref']:g} A accepted and ignored"),
        (p["min_tilt_score"] != 0 or p["max_tilt_score"] != -1,
         f"tilt-particle score window {p['min_tilt_score']:g}..{p['max_tilt_score']:g} accepted and ignored: OCCUPANCY > 0 selects the projections (frealign.py:1762-1765)"),
        (p["smoothing"] != 1, f"smoothing factor {p['smoothing']:g} accepted and ignored"),
        (p["adjust_scores"], "adjust scores for defocus dependence = yes accepted and ignored: scores are used as given"),
        (p["exclude_edges"], "exclude images with blank edges = yes accepted and ignored"),
        (p["crop"], "crop particle images = yes accepted and ignored: images are inserted at their full box"),
        (not p["split_even_odd"], "FSC with even/odd particles = no accepted and ignored: the halves are always odd / even POSITION_IN_STACK (or PIND)"),
        (p["center_mass"], "center mass = yes accepted and ignored"),
        (p["threshold_input"], "threshold input reconstruction = yes accepted and ignored"),
    ]


def dose_weights(p, rows, rows_all, box):
    """Per-projection {weight, cut radius} pairs of the data-driven dose weighting (prompt 22 + its four extra answers,
    frealign.py:1731-1753) or None.  The external file holds one float per scan-order (tilt) index = the mean SCORE of
    that index, -1 if none (inout/metadata/core.py:3039-3075); pyp passes the placeholder `/scratch/not_provided` when
    no file was configured — the weights are then inferred from the parameter file itself, which is what its own
    `compute_global_weights` would have written.  Law: tables.dose_weight_pairs (oracle/SEMANTICS.md §10).
    Returns (pairs, description)."""
    from .. import tables

    if not p.get("dose_weighting"):
        return None, ""
    path = p.get("dose_weights_file", "")
    source = path
    try:
        w = np.loadtxt(path, ndmin=1)
    except (OSError, ValueError):
        w = tables.global_weights(rows_all)
        source = f"the parameter file itself ({path!r} is not a readable weights file)"
    if w.size == 0 or not (w >= 0).any():
        return None, f"dose weighting requested but {source} holds no valid weight: skipped"
    res = p["resolution_limit"] if p["resolution_limit"] > 0 else 2.0 * p["pixel_size"]
    r_rec = min(box * p["pixel_size"] / res, box / 2 - 1)
    pairs = tables.dose_weight_pairs(rows["tind"], w, p.get("dose_fraction", 4), p.get("dose_transition", 0.75), p.get("dose_multiply", True), r_rec)
    n_valid = int((w >= 0).sum())
    keep = max(1, int(np.ceil(n_valid / max(1.0, float(p.get("dose_fraction", 4))))))
    return pairs, (f"Dose weighting from {source}: {n_valid} scan-order indices, the best {keep} at full resolution, the others "
                   f"low-passed at {p.get('dose_transition', 0.75):g} x {r_rec:.1f} Fourier pixels, weights "
                   f"{'scaled to mean 1' if p.get('dose_multiply', True) else 'normalised to sum 1'}")


def run(p, out=sys.stdout, session=None):
    from ..engine import Engine
    from .session import Session

    session = session or Session()
    t0 = time.time()
    hdr = mrc.read_header(p["stack"])
    box = hdr["nx"]
    first, last = p["first"], (min(p["last"], hdr["nz"]) if p["last"] > 0 else hdr["nz"])
    if first < 1 or first > last:
        raise ValueError(f"particle range {first}..{last} outside the stack (1..{hdr['nz']})")
    rows_all = cistem.read_parameters(p["parameters"])
    sel = select_rows(rows_all, first, last)
    rows = rows_all[sel].copy()
    dw, dw_note = dose_weights(p, rows, rows_all, box)
    if dw is not None:
        rows["occupancy"] = np.where(dw[:, 0] > 0, rows["occupancy"], 0)  # indices without a weight are not inserted
    eng = session.engine(first, last - first + 1)
    cfg = Engine.recon_defaults(box, p["pixel_size"])
    cfg.pad = 2 if p["padding"] >= 1.5 else 1
    cfg.mask_radius = p["outer_mask_radius"]
    cfg.resolution_limit = p["resolution_limit"]
    cfg.score_bfactor = p["score_bfactor"]
    cfg.score_weighting = int(p["score_weighting"])
    cfg.score_threshold = p["score_threshold"]
    cfg.normalize, cfg.invert_contrast = int(p["normalize"]), int(p["invert"])
    cfg.per_particle_split = int(p["per_particle_split"])
    used = rows[rows["occupancy"] > 0]
    cfg.average_score = float(used["score"].mean()) if used.size else 0.0
    eng.set_symmetry(p["symmetry"])
    eng.recon_begin(cfg)
    n_band = 0
    if p["likelihood_blurring"] and rows.size:
        # answer 34: the fan of in-plane rotations is scored against the input reconstruction (answer 4)
        from .. import blur

        rcfg = Engine.refine_defaults(box, p["pixel_size"])
        rcfg.pad = cfg.pad
        rcfg.mask_radius = p["outer_mask_radius"]
        rcfg.high_res_limit = p["resolution_limit"] if p["resolution_limit"] > 0 else 2.0 * p["pixel_size"]
        rcfg.normalize, rcfg.invert_contrast = int(p["normalize"]), int(p["invert"])
        eng.ensure_reference(rcfg, p["reference"], lambda: mrc.read(p["reference"])[1])
        n_band = eng.band_counts()[0]
    if rows.size:
        pos = rows["position_in_stack"].astype(np.int64)
        chunk = 8192
        for s in range(0, rows.size, chunk):
            e = min(rows.size, s + chunk)
            imgs = session.images(p["stack"], pos[s:e])
            if n_band:
                blur.insert_blurred(eng, imgs, rows[s:e], n_band, weight_cut=None if dw is None else dw[s:e])
            else:
                eng.recon_insert(imgs, rows[s:e], None if dw is None else dw[s:e])
    n_used = int(used.size)
    if p["dump"]:
        for h, path in ((0, p["dump1"]), (1, p["dump2"])):
            dump.write(path, eng.recon_get_dump(h), box, cfg.pad, h, p["pixel_size"], n_used)
    else:
        from ..formats import statistics

        vol, h1, h2, st = eng.recon_finalize(p["molecular_mass"], p["outer_mask_radius"])
        mrc.write(p["out_map1"], h1, p["pixel_size"])
        mrc.write(p["out_map2"], h2, p["pixel_size"])
        mrc.write(p["out_filtered"], vol, p["pixel_size"])
        with open(p["out_statistics"], "w") as f:
            f.write(statistics.HEADER + statistics.format_table(st) + "\n")
    out.write(banner("Reconstruct3D"))
    out.write(f"\nInserted {n_used} of {rows.size} particles ({first}..{last}), symmetry {p['symmetry']}, "
              f"box {box}, padding {cfg.pad}, {time.time() - t0:.2f} s\n")
    if dw_note:
        out.write(dw_note + "\n")
    if n_band:
        out.write(f"Likelihood blurring: {blur.LBLUR_NROT} in-plane rotations from {blur.LBLUR_START:+.0f} deg in steps of "
                  f"{blur.LBLUR_STEP:.0f} deg, LogP range {blur.LBLUR_RANGE:.0f}\n")
    write_notes(out, "reconstruct3d", ignored_answers(p))
    out.write("\nReconstruct3D: Normal termination\n")
    eng.recon_end()
    session.release()


def main(argv=None):
    try:
        run(parse(Answers(program="reconstruct3d")))
    except (PromptError, ValueError, OSError, RuntimeError, ImportError) as e:
        # pyp greps the log for the word "caught" (src/pyp/refine/csp/particle_cspt.py:812-818)
        sys.stderr.write(f"reconstruct3d: caught error: {e}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
