"""Resolution statistics files exchanged between merge3d and refine3d
(`<name>_statistics.txt`, copied to `statistics_rNN.txt`, src/pyp_main.py:3353-3370) and the
table pyp parses out of the merge3d log (src/pyp/refine/frealign/frealign.py:2558-2567:
fixed widths [5, 8, 10, 10, 10, 10, 10] between the tokens `Rec_SSNR` and
`Merge3D: Normal termination`; column 1 = resolution, column 3 = FSC)."""
import io

import numpy as np

HEADER = ("C                                             Sqrt      Sqrt\n"
          "C NO.   RESOL  RING RAD       FSC  Part_FSC Part_SSNR  Rec_SSNR\n")


def format_table(stats):
    """stats: (n_shells, 7) rows {shell, resolution, ring radius, FSC, Part_FSC, sqrt Part_SSNR,
    sqrt Rec_SSNR}; shell 0 (infinite resolution) is skipped like cisTEM does."""
    lines = []
    for r in np.asarray(stats)[1:]:
        lines.append("%5d%8.2f%10.4f%10.4f%10.4f%10.4f%10.4f" % (int(r[0]), r[1], r[2], r[3], r[4], min(r[5], 9999.0), min(r[6], 9999.0)))
    return "\n".join(lines)


def merge3d_log(stats):
    """Log text whose tail satisfies frealign.py:2558-2567's slicing arithmetic exactly:
    A.find("Rec_SSNR") + 9 is the first table byte and find("Merge3D: ...") - 3 the last."""
    return HEADER + format_table(stats) + "\n\n\nMerge3D: Normal termination\n"


def parse_merge3d_log(text):
    """The reference's own parsing (frealign.py:2558-2567), used by the tests."""
    a = text[text.find("Rec_SSNR") + 9: text.find("Merge3D: Normal termination") - 3]
    widths = [5, 8, 10, 10, 10, 10, 10]
    rows = len(a.split("\n"))
    return np.genfromtxt(io.StringIO(a), delimiter=widths).reshape((rows, len(widths)))


def write_statistics(path, stats):
    """7 x %14.5f per shell, the layout pyp rewrites these files in (postprocess/core.py:219-221)."""
    np.savetxt(path, np.asarray(stats)[1:], fmt="%14.5f%14.5f%14.5f%14.5f%14.5f%14.5f%14.5f")


def read_statistics(path):
    rows = []
    with open(path) as f:
        for line in f:
            t = line.strip()
            if not t or t.startswith("C"):
                t = t[1:].strip() if t.startswith("C") else t
                if not t or not t[0].isdigit():
                    continue
            try:
                rows.append([float(x) for x in t.split()][:7])
            except ValueError:
                continue
    return np.array([r for r in rows if len(r) == 7], dtype=np.float64).reshape(-1, 7)


def ring_weights_from_statistics(stats, box, pixel_size):
    """refine3d 'use statistics' (prompt 6): per-ring SSNR weights w = sqrt(pssnr / (1 + pssnr))
    interpolated onto the box's Fourier rings *(weight law: oracle/SEMANTICS.md)*."""
    stats = np.asarray(stats, dtype=np.float64)
    res, pssnr = stats[:, 1], stats[:, 5] ** 2
    ok = res > 0
    freq = 1.0 / res[ok]
    order = np.argsort(freq)
    rings = np.arange(box + 1, dtype=np.float64) / (box * pixel_size)
    p = np.interp(rings, freq[order], pssnr[ok][order])
    return np.sqrt(p / (1.0 + p)).astype(np.float32)
