"""MRC2000 stacks / volumes as pyp writes and reads them (src/pyp/inout/image/mrc.py:113-156
header fields, :374-435 defaults, :537-559 write): 1024-byte header,
mode 2 = float32, x fastest, nz = number of projections, nsymbt = 0."""
import numpy as np

_MODES = {0: "<i1", 1: "<i2", 2: "<f4", 6: "<u2"}
BYTEORDER_LITTLE = 0x44440000


def read_header(path):
    with open(path, "rb") as f:
        b = f.read(1024)
    if len(b) < 1024:
        raise ValueError(f"{path}: not an MRC file")
    i4 = np.frombuffer(b, dtype="<i4")
    f4 = np.frombuffer(b, dtype="<f4")
    h = {"nx": int(i4[0]), "ny": int(i4[1]), "nz": int(i4[2]), "mode": int(i4[3]), "mx": int(i4[7]), "my": int(i4[8]), "mz": int(i4[9]),
         "xlen": float(f4[10]), "ylen": float(f4[11]), "zlen": float(f4[12]), "amin": float(f4[19]), "amax": float(f4[20]),
         "amean": float(f4[21]), "nsymbt": int(i4[23]), "rms": float(f4[54])}
    if h["mode"] not in _MODES or min(h["nx"], h["ny"], h["nz"]) <= 0:
        raise ValueError(f"{path}: unsupported MRC header (mode {h['mode']}, dims {h['nx']}x{h['ny']}x{h['nz']})")
    h["pixel_size"] = h["xlen"] / h["mx"] if h["mx"] > 0 and h["xlen"] > 0 else 1.0
    return h


def read(path, first=None, last=None, mmap=True):
    """(header, data[z, y, x]); `first`/`last` are 1-based inclusive slice numbers like the
    binaries' prompts (src/pyp/system/local_run.py:513-516).  float32 data is memory-mapped."""
    h = read_header(path)
    nx, ny, nz = h["nx"], h["ny"], h["nz"]
    dt = np.dtype(_MODES[h["mode"]])
    z0 = 0 if first is None else int(first) - 1
    z1 = nz if last is None else int(last)
    if not (0 <= z0 < z1 <= nz):
        raise ValueError(f"{path}: slice range {first}..{last} outside 1..{nz}")
    off = 1024 + h["nsymbt"] + z0 * nx * ny * dt.itemsize
    shape = (z1 - z0, ny, nx)
    if mmap:
        data = np.memmap(path, dtype=dt, mode="r", offset=off, shape=shape)
    else:
        data = np.fromfile(path, dtype=dt, count=int(np.prod(shape)), offset=off).reshape(shape)
    if dt != np.dtype("<f4"):
        data = np.asarray(data, dtype=np.float32)
    return h, data


def make_header(shape, pixel_size=None, stats=None):
    nz, ny, nx = shape
    b = np.zeros(256, dtype="<i4")
    f = b.view("<f4")
    b[0:4] = (nx, ny, nz, 2)
    b[7:10] = (nx, ny, nz)
    px = pixel_size if pixel_size else 1.0
    f[10:13] = (nx * px, ny * px, nz * px)
    f[13:16] = 90.0
    b[16:19] = (1, 2, 3)
    if stats is not None:
        f[19:22] = stats[:3]
        f[54] = stats[3]
    raw = bytearray(b.tobytes())
    raw[208:212] = b"MAP\x00"  # mrc.py strips the trailing blank of "MAP "
    raw[212:216] = np.array([BYTEORDER_LITTLE], dtype="<i4").tobytes()
    return bytes(raw)


def write(path, data, pixel_size=None):
    """Write a float32 volume or stack with the header pyp's mrc.write produces."""
    a = np.ascontiguousarray(data, dtype=np.float32)
    if a.ndim == 2:
        a = a[None]
    stats = (float(a.min()), float(a.max()), float(a.mean(dtype=np.float64)), float(a.std(dtype=np.float64)))
    with open(path, "wb") as f:
        f.write(make_header(a.shape, pixel_size, stats))
        f.write(a.tobytes())


def append(path, data):
    """append_stacks / mrc.merge_fast semantics (mrc.py:643-696): add slices, rewrite nz/mz."""
    a = np.ascontiguousarray(data, dtype=np.float32)
    if a.ndim == 2:
        a = a[None]
    h = read_header(path)
    if (a.shape[2], a.shape[1]) != (h["nx"], h["ny"]) or h["mode"] != 2:
        raise ValueError("Error: can't append stacks with different dimensions")
    with open(path, "r+b") as f:
        f.seek(0, 2)
        f.write(a.tobytes())
        nz = h["nz"] + a.shape[0]
        f.seek(8)
        f.write(np.array([nz], dtype="<i4").tobytes())
        f.seek(36)
        f.write(np.array([nz], dtype="<i4").tobytes())
