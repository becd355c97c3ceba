"""cisTEM `.star` projection tables as the refine_ctf front-end writes them and pyp merges them
(src/pyp/inout/metadata/cistem_star_file.py:1398-1465 `merge_star` / `read_star`): comment lines,
`data_`, `loop_`, one `_cisTEM<Column> #k` line per column, then whitespace separated rows.
Column order = the 32 columns of the binary table (cistem_star_file.py:1404-1437)."""
import numpy as np

from .._lib import ROW_DTYPE

STAR_COLUMNS = [
    "cisTEMPositionInStack", "cisTEMAnglePsi", "cisTEMAngleTheta", "cisTEMAnglePhi", "cisTEMXShift", "cisTEMYShift",
    "cisTEMDefocus1", "cisTEMDefocus2", "cisTEMDefocusAngle", "cisTEMPhaseShift", "cisTEMImageActivity", "cisTEMOccupancy",
    "cisTEMLogP", "cisTEMSigma", "cisTEMScore", "cisTEMPixelSize", "cisTEMMicroscopeVoltagekV", "cisTEMMicroscopeCsMM",
    "cisTEMAmplitudeContrast", "cisTEMBeamTiltX", "cisTEMBeamTiltY", "cisTEMImageShiftX", "cisTEMImageShiftY",
    "cisTEMOriginalXPosition", "cisTEMOriginalYPosition", "cisTEMImageIndex", "cisTEMParticleIndex", "cisTEMTiltIndex",
    "cisTEMRegionIndex", "cisTEMFrameIndex", "cisTEMFrameShiftX", "cisTEMFrameShiftY",
]
assert len(STAR_COLUMNS) == len(ROW_DTYPE.names)


def write_star(path, rows, comment="written by cspb200 refine_ctf"):
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    with open(path, "w") as f:
        f.write(f"# {comment}\n\ndata_\n\nloop_\n")
        for k, name in enumerate(STAR_COLUMNS, start=1):
            f.write(f"_{name} #{k}\n")
        f.write("#    " + " ".join(n[6:10] for n in STAR_COLUMNS) + "\n")
        for r in rows:
            vals = []
            for name in ROW_DTYPE.names:
                v = r[name]
                vals.append(str(int(v)) if ROW_DTYPE[name].kind in "iu" else repr(float(np.float32(v))))
            f.write(" ".join(vals) + "\n")


def read_star(path):
    """Inverse of write_star; also reads files with a column subset / different order."""
    cols, data = [], []
    with open(path) as f:
        lines = iter(f)
        for line in lines:
            if line.strip().lower() == "data_":
                break
        for line in lines:
            s = line.strip()
            if s.startswith("_"):
                cols.append(s.split()[0][1:])
            elif cols and s and not s.startswith("#") and s != "loop_":
                data.append(s.split())
    out = np.zeros(len(data), dtype=ROW_DTYPE)
    for k, c in enumerate(cols):
        if c in STAR_COLUMNS:
            name = ROW_DTYPE.names[STAR_COLUMNS.index(c)]
            out[name] = np.array([float(d[k]) for d in data]).astype(ROW_DTYPE[name])
    return out
