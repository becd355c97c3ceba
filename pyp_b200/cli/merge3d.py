"""merge3d drop-in (src/pyp/refine/frealign/frealign.py:2075-2093; Appendix A.5): answers =
half-map 1, half-map 2, filtered map, statistics file, molecular mass, inner radius, outer radius,
dump seed 1, dump seed 2, N.  Sums the dumps, computes the resolution statistics, applies the
optimal filter, inverse-transforms, corrects the gridding and writes the three maps.  The log
ends with the table pyp parses (frealign.py:2558-2567)."""
import sys
import time

from ..formats import dump, mrc, statistics
from .local_merge3d import sum_dumps
from .prompts import Answers, PromptError, banner, pick_device


def parse(ans: Answers):
    return {"half1": ans.text("output reconstruction 1"), "half2": ans.text("output reconstruction 2"),
            "filtered": ans.text("output filtered reconstruction"), "statistics": ans.text("output resolution statistics"),
            "molecular_mass": ans.number("molecular mass of particle (kDa)"), "inner_radius": ans.number("inner mask radius"),
            "outer_radius": ans.number("outer mask radius"), "seed1": ans.text("input dump seed 1"),
            "seed2": ans.text("input dump seed 2"), "count": ans.integer("number of dump files")}


def run(p, out=sys.stdout, session=None):
    from .session import Session

    session = session or Session()

    if p["count"] < 1:
        raise ValueError("need at least one dump file")
    t = [time.perf_counter()]
    eng = session.engine()
    t.append(time.perf_counter())
    meta, total = sum_dumps(eng, dump.seed_paths(p["seed1"], p["count"]), dump.seed_paths(p["seed2"], p["count"]))
    t.append(time.perf_counter())
    vol, h1, h2, st = eng.recon_finalize(p["molecular_mass"], p["outer_radius"])
    t.append(time.perf_counter())
    mrc.write(p["half1"], h1, meta["pixel_size"])
    mrc.write(p["half2"], h2, meta["pixel_size"])
    mrc.write(p["filtered"], vol, meta["pixel_size"])
    with open(p["statistics"], "w") as f:
        f.write(statistics.HEADER + statistics.format_table(st) + "\n")
    t.append(time.perf_counter())
    out.write(banner("Merge3D"))
    out.write(f"\nMerged {p['count']} dump pairs, {total} particles, box {meta['box']}, pixel {meta['pixel_size']}\n")
    out.write(f"timing: context {t[1] - t[0]:.2f} s, read + sum dumps {t[2] - t[1]:.2f} s, finalise {t[3] - t[2]:.2f} s, write maps {t[4] - t[3]:.2f} s\n\n")
    out.write(statistics.merge3d_log(st))
    eng.recon_end()
    session.release()
    return st


def main(argv=None):
    try:
        run(parse(Answers(program="merge3d")))
    except (PromptError, ValueError, OSError, RuntimeError, ImportError) as e:
        sys.stderr.write(f"merge3d: caught error: {e}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
