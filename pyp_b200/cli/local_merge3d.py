"""local_merge3d drop-in (src/pyp/refine/frealign/frealign.py:1878-1888; Appendix A.4):
answers = output dump 1, output dump 2, seed 1 (`temp_map1_n.mrc`), seed 2, N.  Sums N dump
pairs into one.  The sum runs on the GPU accumulators (cspb_recon_add_dump)."""
import sys

from ..formats import dump
from .prompts import Answers, PromptError, banner, pick_device


def parse(ans: Answers):
    return {"out1": ans.text("output dump 1"), "out2": ans.text("output dump 2"), "seed1": ans.text("input dump seed 1"),
            "seed2": ans.text("input dump seed 2"), "count": ans.integer("number of dump files")}


def sum_dumps(eng, paths1, paths2):
    """Load the dump pairs into a fresh accumulator pair; returns (meta, total inserted)."""
    from ..engine import Engine

    meta = None
    total = 0
    for k, (a, b) in enumerate(zip(paths1, paths2)):
        m1, acc1 = dump.read(a)
        m2, acc2 = dump.read(b)
        if meta is None:
            meta = m1
            cfg = Engine.recon_defaults(m1["box"], m1["pixel_size"])
            cfg.pad = m1["pad"]
            eng.recon_begin(cfg)
        if (m1["box"], m1["pad"]) != (meta["box"], meta["pad"]) or (m2["box"], m2["pad"]) != (meta["box"], meta["pad"]):
            raise ValueError(f"dump {a} does not match the first dump's geometry")
        eng.recon_add_dump(0, acc1)
        eng.recon_add_dump(1, acc2)
        total += m1["n_inserted"]
    return meta, total


def run(p, out=sys.stdout, session=None):
    from .session import Session

    session = session or Session()

    if p["count"] < 1:
        raise ValueError("need at least one dump file")
    eng = session.engine()
    meta, total = sum_dumps(eng, dump.seed_paths(p["seed1"], p["count"]), dump.seed_paths(p["seed2"], p["count"]))
    for h, path in ((0, p["out1"]), (1, p["out2"])):
        dump.write(path, eng.recon_get_dump(h), meta["box"], meta["pad"], h, meta["pixel_size"], total)
    out.write(banner("LocalMerge3D"))
    out.write(f"\nMerged {p['count']} dump pairs ({total} particles)\n\nLocalMerge3D: Normal termination\n")
    eng.recon_end()
    session.release()


def main(argv=None):
    try:
        run(parse(Answers(program="local_merge3d")))
    except (PromptError, ValueError, OSError, RuntimeError, ImportError) as e:
        sys.stderr.write(f"local_merge3d: caught error: {e}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
