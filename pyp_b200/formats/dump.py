"""Intermediate reconstruction dump files (`<name>_map{1,2}_n<k>.mrc`): what reconstruct3d
writes with "dump intermediate arrays = yes" and local_merge3d / merge3d sum
(src/pyp/refine/frealign/frealign.py:1820-1822, 1870-1888, 2087-2091).  pyp never looks inside
them (it renames, tars and moves them), so the payload layout is ours: a 64-byte header followed
by the raw accumulator, one float4 {sum re, sum im, sum ctf^2 w, 0} per voxel of the Hermitian
half-volume [z][y][x], x in [0, np/2], y and z centred."""
import numpy as np

MAGIC = b"CSPBDUMP"
_HDR = np.dtype([("magic", "S8"), ("version", "<i4"), ("box", "<i4"), ("pad", "<i4"), ("half", "<i4"),
                 ("pixel_size", "<f4"), ("n_inserted", "<i8"), ("reserved", "S28")])
assert _HDR.itemsize == 64


def write(path, acc, box, pad, half, pixel_size, n_inserted):
    npad = box * pad
    acc = np.ascontiguousarray(acc, dtype=np.float32)
    assert acc.size == npad * npad * (npad // 2 + 1) * 4
    h = np.zeros(1, dtype=_HDR)
    h["magic"], h["version"], h["box"], h["pad"], h["half"] = MAGIC, 1, box, pad, half
    h["pixel_size"], h["n_inserted"] = pixel_size, n_inserted
    with open(path, "wb") as f:
        f.write(h.tobytes())
        f.write(acc.tobytes())


def read(path):
    with open(path, "rb") as f:
        h = np.frombuffer(f.read(64), dtype=_HDR)[0]
        if h["magic"] != MAGIC:
            raise ValueError(f"{path}: not a cspb200 reconstruction dump")
        npad = int(h["box"]) * int(h["pad"])
        n = npad * npad * (npad // 2 + 1) * 4
        acc = np.fromfile(f, dtype="<f4", count=n)
    if acc.size != n:
        raise ValueError(f"{path}: truncated dump")
    meta = {"box": int(h["box"]), "pad": int(h["pad"]), "half": int(h["half"]), "pixel_size": float(h["pixel_size"]), "n_inserted": int(h["n_inserted"])}
    return meta, acc.reshape(npad, npad, npad // 2 + 1, 4)


def seed_paths(seed, count):
    """`<stem>_n.mrc` -> [`<stem>_n1.mrc` ... `<stem>_nN.mrc`] (frealign.py:1870-1885, 2087-2091)."""
    if not seed.endswith("_n.mrc"):
        raise ValueError(f"dump seed {seed!r} must end with _n.mrc")
    stem = seed[: -len(".mrc")]
    return [f"{stem}{k}.mrc" for k in range(1, int(count) + 1)]
