"""refine3d drop-in: same stdin answer list as external/cistem2/refine3d
(src/pyp/refine/frealign/frealign.py:3918-3994), numerics on the GPU.

Reads: particle stack (MRC), input `.cistem`, optional `_stat.cistem`, reference map, statistics.
Writes: `<out>.cistem` with the rows of [first, last] (PSI, THETA, PHI, X_SHIFT, Y_SHIFT, LOGP,
SIGMA, SCORE updated) and `<out>_changes.cistem`.  Exit status 0 on success, non-zero with the
word `caught` on stderr otherwise (pyp only checks the output files, frealign.py:3086-3094).
"""
import os
import sys
import time

import numpy as np

from ..formats import cistem, mrc, statistics
from .prompts import Answers, PromptError, banner, pick_device, write_notes


def parse(ans: Answers):
    p = {}
    p["stack"] = ans.text("input particle images")
    p["parameters"] = ans.text("input cisTEM parameter file")
    p["global_stat"] = ans.text("global statistics parameter file")
    p["reference"] = ans.text("input reconstruction")
    p["statistics"] = ans.text("input data statistics")
    p["use_statistics"] = ans.yesno("use statistics")
    p["use_priors"] = ans.yesno("use priors")
    p["matching_out"] = ans.text("output matching projections")
    p["out_parameters"] = ans.text("output parameter file")
    p["out_changes"] = ans.text("output parameter changes")
    p["symmetry"] = ans.text("particle symmetry")
    p["first"] = ans.integer("first particle to refine")
    p["last"] = ans.integer("last particle to refine")
    p["percent_used"] = ans.number("percent of particles to use")
    p["pixel_size"] = ans.number("pixel size of reconstruction")
    p["molecular_mass"] = ans.number("molecular mass of particle (kDa)")
    p["inner_mask_radius"] = ans.number("inner mask radius")
    p["outer_mask_radius"] = ans.number("outer mask radius")
    p["low_res_limit"] = ans.number("low resolution limit")
    p["high_res_limit"] = ans.number("high resolution limit")
    p["signed_cc_limit"] = ans.number("resolution limit for signed CC")
    p["class_res_limit"] = ans.number("resolution limit for classification")
    p["search_mask_radius"] = ans.number("mask radius for global search")
    p["search_high_res"] = ans.number("approx. resolution limit for search")
    p["angular_step"] = ans.number("angular step")
    p["best_matches"] = ans.integer("number of top hits to refine")
    p["search_range_x"] = ans.number("search range in X")
    p["search_range_y"] = ans.number("search range in Y")
    p["mask_2d"] = [ans.number(f"2D mask {k}") for k in ("X", "Y", "Z", "radius")]
    p["defocus_range"] = ans.number("defocus search range")
    p["defocus_step"] = ans.number("defocus step")
    p["padding"] = ans.number("tuning parameter: padding factor")
    p["global_search"] = ans.yesno("global search")
    p["local_refine"] = ans.yesno("local refinement")
    for k in ("psi", "theta", "phi", "x", "y"):
        p[f"refine_{k}"] = ans.yesno(f"refine {k}")
    p["calc_matching"] = ans.yesno("calculate matching projections")
    p["apply_2d_masking"] = ans.yesno("apply 2D masking")
    p["refine_defocus"] = ans.yesno("refine defocus")
    p["normalize"] = ans.yesno("normalize particles")
    p["invert"] = ans.yesno("invert particle contrast")
    p["exclude_edges"] = ans.yesno("exclude images with blank edges")
    p["normalize_rec"] = ans.yesno("normalize input reconstruction")
    p["threshold_rec"] = ans.yesno("threshold input reconstruction")
    return p


def select_rows(rows, first, last):
    """Rows whose POSITION_IN_STACK lies in the 1-based inclusive range (local_run.py:513-516)."""
    pos = rows["position_in_stack"].astype(np.int64)
    sel = np.nonzero((pos >= first) & (pos <= last))[0]
    return sel[np.argsort(pos[sel], kind="stable")]


def shift_prior(p, rows_all):
    """Mean / variance of X_SHIFT, Y_SHIFT (Angstrom) for the shift restraint of answer 7: rows 0 / 1
    of the global statistics file (`<name>_stat.cistem` = np.mean / np.var of the merged used rows,
    src/pyp/refine/csp/particle_cspt.py:1009-1016; handed over at frealign.py:3827-3831) when it
    exists, else the same statistics of the input parameter file."""
    stat = p["global_stat"]
    if stat and stat != "null" and os.path.exists(stat):
        st = cistem.read_parameters(stat)
        if st.size >= 2:
            return (float(st["x_shift"][0]), float(st["y_shift"][0]), float(st["x_shift"][1]), float(st["y_shift"][1]))
    if rows_all.size == 0:
        return 0.0, 0.0, 0.0, 0.0
    x, y = rows_all["x_shift"].astype(np.float64), rows_all["y_shift"].astype(np.float64)
    return float(x.mean()), float(y.mean()), float(x.var()), float(y.var())


def ignored_answers(p):
    """Answers of the list that this implementation reads and does not act on (each gets a log line)."""
    return [
        (abs(p["percent_used"] - 1.0) > 1e-6, f"percent of particles to use = {p['percent_used']:g} accepted and ignored: every particle of the range is refined"),
        (p["inner_mask_radius"] != 0, f"inner mask radius {p['inner_mask_radius']:g} A accepted and ignored (pyp always passes 0)"),
        (p["class_res_limit"] > 0 and abs(p["class_res_limit"] - p["high_res_limit"]) > 1e-6,
         f"resolution limit for classification {p['class_res_limit']:g} A accepted and ignored: LOGP is computed on the scoring band"),
        (p["global_search"], f"mask radius for global search {p['search_mask_radius']:g} A accepted and ignored: the search uses the outer mask radius"),
        (p["exclude_edges"], "exclude images with blank edges = yes accepted and ignored"),
        (p["normalize_rec"], "normalize input reconstruction = yes accepted and ignored: the reference is used as given"),
        (p["threshold_rec"], "threshold input reconstruction = yes accepted and ignored: the reference is used as given"),
    ]


def build_cfg(p, box):
    from ..engine import Engine

    cfg = Engine.refine_defaults(box, p["pixel_size"])
    cfg.pad = 2 if p["padding"] >= 1.5 else 1
    cfg.mask_radius = p["outer_mask_radius"]
    cfg.low_res_limit = p["low_res_limit"]
    cfg.high_res_limit = p["high_res_limit"]
    cfg.signed_cc_limit = p["signed_cc_limit"]
    cfg.search_mask_radius = p["search_mask_radius"]
    cfg.search_high_res = p["search_high_res"]
    cfg.angular_step = p["angular_step"]
    cfg.best_matches = p["best_matches"]
    cfg.search_range_x, cfg.search_range_y = p["search_range_x"], p["search_range_y"]
    cfg.defocus_range, cfg.defocus_step = p["defocus_range"], p["defocus_step"]
    cfg.global_search, cfg.local_refine = int(p["global_search"]), int(p["local_refine"])
    cfg.refine_psi, cfg.refine_theta, cfg.refine_phi = int(p["refine_psi"]), int(p["refine_theta"]), int(p["refine_phi"])
    cfg.refine_x, cfg.refine_y = int(p["refine_x"]), int(p["refine_y"])
    cfg.refine_defocus = int(p["refine_defocus"])
    cfg.normalize, cfg.invert_contrast = int(p["normalize"]), int(p["invert"])
    return cfg


def run(p, out=sys.stdout, session=None):
    from .session import Session

    session = session or Session()
    t0 = time.time()
    hdr = mrc.read_header(p["stack"])
    box = hdr["nx"]
    if hdr["ny"] != box:
        raise ValueError("particle images must be square")
    first, last = p["first"], min(p["last"], hdr["nz"]) if p["last"] > 0 else hdr["nz"]
    if first < 1 or first > last:
        raise ValueError(f"particle range {first}..{last} outside the stack (1..{hdr['nz']})")
    rows_all = cistem.read_parameters(p["parameters"])
    sel = select_rows(rows_all, first, last)
    rows = rows_all[sel]
    eng = session.engine(first, last - first + 1)
    cfg = build_cfg(p, box)
    if p["use_priors"]:
        cfg.use_priors = 1
        cfg.prior_mean_x, cfg.prior_mean_y, cfg.prior_var_x, cfg.prior_var_y = shift_prior(p, rows_all)
    reused = eng.ensure_reference(cfg, p["reference"], lambda: mrc.read(p["reference"])[1])
    if p["use_statistics"] and os.path.exists(p["statistics"]) and os.path.getsize(p["statistics"]) > 0:
        st = statistics.read_statistics(p["statistics"])
        if st.size:
            eng.set_ring_weights(statistics.ring_weights_from_statistics(st, box, p["pixel_size"]))
    focus = p["apply_2d_masking"] and p["mask_2d"][3] > 0
    eng.set_focus_mask(*(p["mask_2d"] if focus else (0, 0, 0, 0)))
    eng.set_symmetry(p["symmetry"])
    if p["global_search"]:
        from ..search_grid import search_grid

        eng.set_search_grid(search_grid(p["angular_step"], p["symmetry"]))
    pos = rows["position_in_stack"].astype(np.int64)
    chunk = 16384
    for s in range(0, rows.size, chunk):
        eng.load_images(session.images(p["stack"], pos[s:s + chunk]), append=s > 0)
    refined, changes, n_evals = eng.refine(rows, want_changes=True) if rows.size else (rows, rows.copy(), 0)
    if rows.size:  # no zero-row files: the reference's reader rejects an empty payload (cistem_star_file.py:694-776)
        cistem.write_parameters(p["out_parameters"], refined)
        cistem.write_parameters(p["out_changes"], changes)
        if p["calc_matching"]:  # answers 8 / 43: `<name>_match.mrc_<range>`, one matching projection per refined row
            mrc.write(p["matching_out"], eng.matching_projections(refined), p["pixel_size"])
    dt = time.time() - t0
    out.write(banner("Refine3D"))
    out.write(f"\nRefining particles {first} to {last} ({rows.size} rows), box {box}, pixel {p['pixel_size']}\n")
    if rows.size:
        out.write(f"Mean score {float(refined['score'].mean()):.4f}, mean change {float(changes['score'].mean()):+.4f}, "
                  f"{n_evals} projections scored in {dt:.2f} s\n")
    if p["use_priors"]:
        out.write(f"Shift restraint: mean ({cfg.prior_mean_x:.3f}, {cfg.prior_mean_y:.3f}) A, "
                  f"variance ({cfg.prior_var_x:.3f}, {cfg.prior_var_y:.3f}) A^2\n")
    if focus:
        out.write("LogP evaluated inside the 2-D focus mask: centre ({:.1f}, {:.1f}, {:.1f}) A, radius {:.1f} A\n".format(*p["mask_2d"]))
    if p["calc_matching"] and rows.size:
        out.write(f"Matching projections written to {p['matching_out']}\n")
    if reused:
        out.write("Reference transform reused from the resident engine\n")
    write_notes(out, "refine3d", ignored_answers(p))
    out.write("\nRefine3D: Normal termination\n")
    session.release()
    return refined


def main(argv=None):
    try:
        run(parse(Answers(program="refine3d")))
    except (PromptError, ValueError, OSError, RuntimeError, ImportError) as e:
        sys.stderr.write(f"refine3d: caught error: {e}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
