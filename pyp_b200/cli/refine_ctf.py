"""refine_ctf drop-in: same stdin answer list as external/cistem2/refine_ctf
(src/pyp/refine/frealign/frealign.py:3995-4041), numerics on the GPU.

Per-particle defocus refinement with the refine3d scorer (pose fixed, defocus offset free).  Writes
`<out>_refined_ctf.star` / `<out>_changes.star` (merged by pyp with merge_star,
frealign.py:3133-3154).  Beam-tilt estimation (answer 23): the GPU sums G * conj(CTF * slice) over
the particles of the range (`cspb_refine_phase_sum`), the coma phase of a tilted beam is fitted to it
(`pyp_b200/beamtilt.py`, oracle/SEMANTICS.md §12); BEAM_TILT_X/Y (mrad) and IMAGE_SHIFT_X/Y (Angstrom)
of every row of the range are set to the fit and the three diagnostic images are written.
"""
import sys
import time

import numpy as np

from ..formats import cistem, mrc, star
from .prompts import Answers, PromptError, banner, pick_device
from .refine3d import select_rows


def parse(ans: Answers):
    p = {}
    p["stack"] = ans.text("input particle images")
    p["parameters"] = ans.text("input cisTEM parameter file")
    p["reference"] = ans.text("input reconstruction")
    p["statistics"] = ans.text("input data statistics")
    p["use_statistics"] = ans.yesno("use statistics")
    p["out_star"] = ans.text("output star file")
    p["out_changes"] = ans.text("output parameter changes")
    p["phase_difference"] = ans.text("output phase difference image")
    p["beamtilt_image"] = ans.text("output beam tilt image")
    p["difference_image"] = ans.text("output difference image")
    p["first"] = ans.integer("first particle to refine")
    p["last"] = ans.integer("last particle to refine")
    p["pixel_size"] = ans.number("pixel size of reconstruction")
    p["molecular_mass"] = ans.number("molecular mass of particle (kDa)")
    p["inner_mask_radius"] = ans.number("inner mask radius")
    p["outer_mask_radius"] = ans.number("outer mask radius")
    p["low_res_limit"] = ans.number("low resolution limit")
    p["high_res_limit"] = ans.number("high resolution limit")
    p["defocus_range"] = ans.number("defocus search range")
    p["defocus_step"] = ans.number("defocus step")
    p["padding"] = ans.number("tuning parameter: padding factor")
    p["refine_defocus"] = ans.yesno("refine defocus")
    p["beam_tilt"] = ans.yesno("estimate beam tilt")
    p["normalize"] = ans.yesno("normalize particles")
    p["invert"] = ans.yesno("invert particle contrast")
    p["exclude_edges"] = ans.yesno("exclude images with blank edges")
    p["normalize_rec"] = ans.yesno("normalize input reconstruction")
    p["threshold_rec"] = ans.yesno("threshold input reconstruction")
    return p


def run(p, out=sys.stdout, session=None):
    from ..engine import Engine
    from .session import Session

    session = session or Session()

    t0 = time.time()
    hdr = mrc.read_header(p["stack"])
    box = hdr["nx"]
    first, last = p["first"], min(p["last"], hdr["nz"]) if p["last"] > 0 else hdr["nz"]
    if first < 1 or first > last:
        raise ValueError(f"particle range {first}..{last} outside the stack (1..{hdr['nz']})")
    rows_all = cistem.read_parameters(p["parameters"])
    rows = rows_all[select_rows(rows_all, first, last)]
    refined, n_evals = rows.copy(), 0
    tilt = None
    if (p["refine_defocus"] or p["beam_tilt"]) and rows.size:
        eng = session.engine(first, last - first + 1)
        cfg = Engine.refine_defaults(box, p["pixel_size"])
        cfg.pad = 2 if p["padding"] >= 1.5 else 1
        cfg.mask_radius = p["outer_mask_radius"]
        cfg.low_res_limit, cfg.high_res_limit = p["low_res_limit"], p["high_res_limit"]
        cfg.defocus_range, cfg.defocus_step = p["defocus_range"], p["defocus_step"]
        cfg.global_search, cfg.local_refine = 0, 1
        cfg.refine_psi = cfg.refine_theta = cfg.refine_phi = cfg.refine_x = cfg.refine_y = 0
        cfg.refine_defocus = int(p["refine_defocus"])
        cfg.normalize, cfg.invert_contrast = int(p["normalize"]), int(p["invert"])
        eng.ensure_reference(cfg, p["reference"], lambda: mrc.read(p["reference"])[1])
        pos = rows["position_in_stack"].astype(np.int64)
        for s in range(0, rows.size, 16384):
            eng.load_images(session.images(p["stack"], pos[s:s + 16384]), append=s > 0)
        if p["refine_defocus"]:
            refined, _, n_evals = eng.refine(rows)
            # keep the search inside +- defocus_range of the input (answer 19)
            if p["defocus_range"] > 0:
                d = np.clip(refined["defocus_1"] - rows["defocus_1"], -p["defocus_range"], p["defocus_range"])
                refined["defocus_1"], refined["defocus_2"] = rows["defocus_1"] + d, rows["defocus_2"] + d
        if p["beam_tilt"]:
            from .. import beamtilt

            S = eng.phase_sum(refined)
            n_evals += int(rows.size)
            tilt = beamtilt.fit(S, p["pixel_size"], float(rows["voltage_kv"][0]), float(rows["cs_mm"][0]))
            refined["beam_tilt_x"], refined["beam_tilt_y"] = tilt["beam_tilt_x"], tilt["beam_tilt_y"]
            refined["image_shift_x"], refined["image_shift_y"] = tilt["shift_x"], tilt["shift_y"]
        session.release()
    changes = refined.copy()
    for k in ("defocus_1", "defocus_2", "score", "logp", "sigma", "beam_tilt_x", "beam_tilt_y", "image_shift_x", "image_shift_y"):
        changes[k] = refined[k] - rows[k]
    star.write_star(p["out_star"], refined)
    star.write_star(p["out_changes"], changes)
    zero = np.zeros((box, box), dtype=np.float32)
    for k, img in (("phase_difference", "phase"), ("beamtilt_image", "model"), ("difference_image", "difference")):
        mrc.write(p[k], (tilt[img] if tilt else zero)[None], p["pixel_size"])
    out.write(banner("RefineCTF"))
    out.write(f"\nRefining defocus of particles {first} to {last} ({rows.size} rows), {n_evals} projections scored in {time.time() - t0:.2f} s\n")
    if rows.size:
        out.write(f"Mean defocus change {float(changes['defocus_1'].mean()):+.1f} A, mean score change {float(changes['score'].mean()):+.4f}\n")
    if tilt:
        out.write(f"Beam tilt ({tilt['beam_tilt_x']:+.4f}, {tilt['beam_tilt_y']:+.4f}) mrad, particle shift "
                  f"({tilt['shift_x']:+.4f}, {tilt['shift_y']:+.4f}) A, weighted phase residual {tilt['rms']:.4f} rad\n")
    out.write("\nRefineCTF: Normal termination\n")
    return refined


def main(argv=None):
    try:
        run(parse(Answers(program="refine_ctf")))
    except (PromptError, ValueError, OSError, RuntimeError, ImportError) as e:
        sys.stderr.write(f"refine_ctf: caught error: {e}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
