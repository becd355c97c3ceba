"""`.cistem` binary parameter tables — bit-compatible with the reference's reader/writer
(src/pyp/inout/metadata/cistem_star_file.py:694-776 projections, :266-382 extended blocks;
layout in SURVEY.md Appendix B).

    Parameters file:  int32 ncols | int32 nrows | ncols x {int64 column_id, int8 type} | rows
    Extended file:    for block in (PIND=15, TIND=35):
                          int64 block_id | int32 ncols | int32 nrows | ncols x {int64, int8} | rows

Unlike the reference (``np.array(data.tolist())``, cistem_star_file.py:723) rows stay a packed
structured array (:data:`ROW_DTYPE`, 128 bytes) that goes to the GPU unchanged.
"""
import numpy as np

from .._lib import PARTICLE_DTYPE, ROW_DTYPE, TILT_DTYPE

# type codes — cistem_star_file.py:18-27
INTEGER, FLOAT, LONG, CHAR, INTEGER_UNSIGNED = 2, 3, 5, 7, 9
_TYPE_NP = {INTEGER: "<i4", FLOAT: "<f4", LONG: "<i8", CHAR: "<i1", INTEGER_UNSIGNED: "<u4"}

# column ids — cistem_star_file.py:30-91
COLUMN_IDS = {
    "position_in_stack": 1, "image_is_active": 2, "psi": 4, "x_shift": 8, "y_shift": 16,
    "defocus_1": 32, "defocus_2": 64, "defocus_angle": 128, "phase_shift": 256, "occupancy": 512,
    "logp": 1024, "sigma": 2048, "score": 4096, "score_change": 8192, "pixel_size": 16384,
    "voltage_kv": 32768, "cs_mm": 65536, "amplitude_contrast": 131072, "beam_tilt_x": 262144,
    "beam_tilt_y": 524288, "image_shift_x": 1048576, "image_shift_y": 2097152, "theta": 4194304,
    "phi": 8388608, "original_x": 8589934592, "original_y": 17179869184,
    "imind": 20, "pind": 15, "tind": 35, "rind": 70, "find": 55, "fshift_x": 11, "fshift_y": 121,
}
_ID_TO_NAME = {v: k for k, v in COLUMN_IDS.items()}
_ROW_TYPES = {name: (INTEGER_UNSIGNED if ROW_DTYPE[name].kind == "u" else INTEGER if ROW_DTYPE[name].kind == "i" else FLOAT) for name in ROW_DTYPE.names}

PIND_BLOCK, TIND_BLOCK = 15, 35
# extended blocks — cistem_star_file.py:247-248
_PARTICLE_IDS = [15, 3, 9, 27, 81, 273, 819, 2457, 7371, 22113, 66339, 199017]
_TILT_IDS = [35, 70, 7, 49, 343, 2401]
_HDR = np.dtype([("id", "<i8"), ("type", "<i1")])


def _read_header(buf, off):
    ncols, nrows = np.frombuffer(buf, dtype="<i4", count=2, offset=off)
    off += 8
    cols = np.frombuffer(buf, dtype=_HDR, count=int(ncols), offset=off)
    off += int(ncols) * 9
    return int(ncols), int(nrows), cols, off


def read_parameters(path):
    """Read a projection table into a packed ROW_DTYPE array (missing columns stay zero)."""
    buf = open(path, "rb").read()
    if len(buf) < 8:
        raise ValueError(f"{path}: binary file is broken")
    ncols, nrows, cols, off = _read_header(buf, 0)
    fields = []
    for cid, tcode in cols:
        cid, tcode = int(cid), int(tcode)
        if cid not in _ID_TO_NAME or tcode not in _TYPE_NP:
            raise ValueError(f"{path}: unrecognised column code {cid} (type {tcode})")
        fields.append((_ID_TO_NAME[cid], _TYPE_NP[tcode]))
    dt = np.dtype(fields)
    if len(buf) - off < nrows * dt.itemsize:
        raise ValueError(f"{path}: truncated ({len(buf) - off} bytes for {nrows} rows of {dt.itemsize})")
    raw = np.frombuffer(buf, dtype=dt, count=nrows, offset=off)
    if dt == ROW_DTYPE:
        return raw.copy()
    out = np.zeros(nrows, dtype=ROW_DTYPE)
    for name in dt.names:
        if name in ROW_DTYPE.names:
            out[name] = raw[name]
    return out


def write_parameters(path, rows):
    """Write the standard 32-column table exactly as Parameters.to_binary does."""
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    with open(path, "wb") as f:
        f.write(np.array([len(ROW_DTYPE.names), rows.size], dtype="<i4").tobytes())
        hdr = np.zeros(len(ROW_DTYPE.names), dtype=_HDR)
        for k, name in enumerate(ROW_DTYPE.names):
            hdr[k] = (COLUMN_IDS[name], _ROW_TYPES[name])
        f.write(hdr.tobytes())
        f.write(rows.tobytes())


def extended_path(path):
    return path.replace(".cistem", "_extended.cistem")


def read_extended(path):
    """Returns (particles PARTICLE_DTYPE array, tilts TILT_DTYPE array)."""
    buf = open(path, "rb").read()
    off = 0
    out = {}
    for _ in range(2):
        block = int(np.frombuffer(buf, dtype="<i8", count=1, offset=off)[0])
        off += 8
        ncols, nrows, cols, off = _read_header(buf, off)
        want_ids, dt = (_PARTICLE_IDS, PARTICLE_DTYPE) if block == PIND_BLOCK else (_TILT_IDS, TILT_DTYPE)
        if block not in (PIND_BLOCK, TIND_BLOCK) or [int(c) for c in cols["id"]] != want_ids:
            raise ValueError(f"{path}: unexpected extended block {block}")
        out[block] = np.frombuffer(buf, dtype=dt, count=nrows, offset=off).copy()
        off += nrows * dt.itemsize
    return out[PIND_BLOCK], out[TIND_BLOCK]


def write_extended(path, particles, tilts):
    with open(path, "wb") as f:
        for block, ids, dt, data in ((PIND_BLOCK, _PARTICLE_IDS, PARTICLE_DTYPE, particles), (TIND_BLOCK, _TILT_IDS, TILT_DTYPE, tilts)):
            data = np.ascontiguousarray(data, dtype=dt)
            f.write(np.array([block], dtype="<i8").tobytes())
            f.write(np.array([len(ids), data.size], dtype="<i4").tobytes())
            hdr = np.zeros(len(ids), dtype=_HDR)
            for k, (cid, name) in enumerate(zip(ids, dt.names)):
                hdr[k] = (cid, INTEGER if dt[name].kind == "i" else FLOAT)
            f.write(hdr.tobytes())
            f.write(data.tobytes())


def merge(paths):
    """Parameters.merge (cistem_star_file.py:656-692): stack and sort by POSITION_IN_STACK."""
    if not paths:
        raise ValueError("No cistem binary file to merge.")
    rows = np.concatenate([read_parameters(p) for p in paths])
    return rows[np.argsort(rows["position_in_stack"], kind="stable")]


def merge_extended(paths):
    """The extended half of Parameters.merge (cistem_star_file.py:674-686): particle and tilt tables of the
    files are overlaid in list order with dict.update semantics — a later file replaces the entries it holds
    (by PIND, by (TIND, RIND)), new entries are appended, the order of first appearance is kept.  pyp passes
    the un-refined table first and the csp outputs after it (particle_cspt.py:122-128)."""
    if not paths:
        raise ValueError("No cistem extended binary file to merge.")

    def overlay(tables, keys):
        order, latest = [], {}
        for t in tables:
            for row in t:
                k = tuple(int(row[name]) for name in keys)
                if k not in latest:
                    order.append(k)
                latest[k] = row
        if not order:
            return tables[0][:0].copy()
        return np.array([latest[k] for k in order], dtype=tables[0].dtype)

    both = [read_extended(p) for p in paths]
    particles = overlay([b[0] for b in both], ("pind",))
    tilts = [b[1] for b in both]
    # tilts: dict of dicts — outer order by first appearance of TIND, inner by first appearance of RIND
    merged = overlay(tilts, ("tind", "rind"))
    if merged.size:
        first_seen = {}
        for k, t in enumerate(merged["tind"]):
            first_seen.setdefault(int(t), k)
        rank = np.array([first_seen[int(t)] for t in merged["tind"]])
        merged = merged[np.argsort(rank, kind="stable")]
    return particles, merged


def merge_with_film_id(paths):
    """merge_all_binary_with_filmid (cistem_star_file.py:1495-1550): stack the per-film tables in list
    order, IMAGE_IS_ACTIVE (which pyp re-uses as the film index) = position of the file in the list.
    Rows stay packed; nothing goes through Python lists."""
    parts = []
    for film, p in enumerate(paths):
        rows = read_parameters(p)
        rows["image_is_active"] = film
        parts.append(rows)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=ROW_DTYPE)
