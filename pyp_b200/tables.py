"""Table bookkeeping pyp does around the `csp` calls, vectorised over the packed rows:

  * ``update_particle_score`` — after the csp processes are merged, every particle's SCORE becomes the
    mean SCORE of its projections inside the exposure window (TIND range) or a tilt-angle range;
    particles with no projection in the window get score -1 and occupancy 0
    (src/pyp/inout/metadata/cistem_star_file.py:936-986; called at src/pyp/align/core.py:1204 and
    src/pyp/analysis/scores.py:929);
  * ``sync_particle_occ`` — copy occupancies between the particle table and the projection rows
    (cistem_star_file.py:988-1013; align/core.py:912, scores.py:975).

  * ``global_weights`` / ``write_global_weights`` — the external dose-weighting file reconstruct3d reads when
    prompt 22 is "yes": one value per scan-order index (TIND) = mean SCORE of the projections with
    OCCUPANCY > 0 at that index, -1 where there is none (src/pyp/inout/metadata/core.py:3039-3075).

  * ``parameter_statistics`` — the two rows of ``<name>_stat.cistem``: column means and variances of the
    merged used rows (src/pyp/refine/csp/particle_cspt.py:1009-1016), which refine3d / reconstruct3d get as
    answer 3 and the shift restraint of refine3d reads (oracle/SEMANTICS.md §7b).

Pinned against the reference's own methods by tests/golden/tables_* (tests/golden/make_golden_tables.py).
"""
import numpy as np


def _groups(keys):
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    starts = np.flatnonzero(np.r_[True, sk[1:] != sk[:-1]])
    return order, sk[starts], starts, np.r_[starts[1:], sk.size]


def update_particle_score(rows, particles, tilts=None, tind_range=(0, -1), tiltang_range=(-90, 90)):
    """Return a copy of `particles` (PARTICLE_DTYPE) with score / occ updated from `rows` (ROW_DTYPE).
    A non-empty `tind_range` (min, max; max = -1: open) wins over `tiltang_range`, as in the reference."""
    use_tind, use_ang = len(tind_range) > 0, len(tiltang_range) > 0
    if not (use_tind or use_ang):
        raise ValueError("give a TIND range or a tilt-angle range")
    out = particles.copy()
    tind = rows["tind"].astype(np.int64)
    if use_tind:
        lo, hi = int(tind_range[0]), int(tind_range[1])
        inside = ~((tind < lo) | ((tind > hi) & (hi != -1)))
    else:
        lo, hi = float(tiltang_range[0]), float(tiltang_range[1])
        if lo > hi:
            raise ValueError(f"min angle ({lo}) should be smaller than max angle ({hi})")
        if tilts is None:
            raise ValueError("a tilt-angle range needs the tilt table")
        key = {(int(t["tind"]), int(t["rind"])): float(t["angle"]) for t in tilts}
        ang = np.array([key[(int(a), int(b))] for a, b in zip(rows["tind"], rows["rind"])], dtype=np.float64)
        inside = (ang >= lo) & (ang <= hi)
    score = rows["score"].astype(np.float64)[inside]
    pind = rows["pind"].astype(np.int64)[inside]
    mean = {}
    if pind.size:
        order, ids, starts, ends = _groups(pind)
        s = score[order]
        for k, a, b in zip(ids, starts, ends):
            mean[int(k)] = np.mean(s[a:b])  # np.mean per particle, as the reference (float64 pairwise sum)
    for i in range(out.size):
        m = mean.get(int(out["pind"][i]))
        if m is None:
            out["score"][i] = -1.0
            out["occ"][i] = 0.0
        else:
            out["score"][i] = m
    return out


def sync_particle_occ(rows, particles, ptl_to_prj=True):
    """ptl_to_prj: projections take their particle's occupancy; else particles take the mean occupancy of
    their projections.  Returns (rows, particles) copies; entries without a partner are left alone."""
    rows, particles = rows.copy(), particles.copy()
    if ptl_to_prj:
        ids = particles["pind"].astype(np.int64)
        order = np.argsort(ids, kind="stable")
        pos = np.searchsorted(ids[order], rows["pind"].astype(np.int64))
        pos = np.clip(pos, 0, max(ids.size - 1, 0))
        if ids.size:
            hit = ids[order][pos] == rows["pind"]
            rows["occupancy"][hit] = particles["occ"][order][pos][hit]
    else:
        pind = rows["pind"].astype(np.int64)
        if pind.size:
            order, gid, starts, ends = _groups(pind)
            occ = rows["occupancy"].astype(np.float64)[order]
            mean = {int(k): np.mean(occ[a:b]) for k, a, b in zip(gid, starts, ends)}
            for i in range(particles.size):
                m = mean.get(int(particles["pind"][i]))
                if m is not None:
                    particles["occ"][i] = m
    return rows, particles


def global_weights(rows):
    """Mean SCORE per scan-order index over the projections with occupancy > 0; -1 for unused indices below
    the largest used one (core.py:3039-3071)."""
    used = rows[rows["occupancy"] > 0.0]
    if used.size == 0:
        return np.zeros(0, dtype=np.float64)
    tind = used["tind"].astype(np.int64)
    total = np.bincount(tind, weights=used["score"].astype(np.float64))
    count = np.bincount(tind)
    out = np.full(total.size, -1.0)
    out[count > 0] = total[count > 0] / count[count > 0]
    return out


def write_global_weights(path, weights):
    """One Python-float repr per line, no trailing newline (core.py:3073-3074)."""
    with open(path, "w") as f:
        f.write("\n".join(str(float(w)) for w in weights))


def dose_weight_pairs(tind, weights, fraction, transition, multiply, r_rec):
    """Per-projection {weight, cut radius} of reconstruct3d's data-driven dose weighting (prompt 22, frealign.py:1731-1753;
    the fork's law is not public — this is oracle/SEMANTICS.md §10).

    `weights` = one value per scan-order index (TIND): the mean SCORE at that index, -1 where unused (`global_weights`).
    The valid weights are normalised to sum 1 and, with `multiply`, scaled by their number ("multiply by number of
    frames"): a flat series gives weight 1 everywhere.  The best ceil(n / fraction) indices contribute at every
    resolution ("larger values contribute fewer frames to high resolution"); the others are low-passed at
    `transition` x `r_rec` (Fourier pixels; the insertion kernel applies a raised-cosine edge there).  Projections
    whose TIND has no valid weight get weight 0.  Returns an (n_rows, 2) float32 array."""
    w = np.asarray(weights, dtype=np.float64).ravel()
    tind = np.asarray(tind, dtype=np.int64)
    out = np.zeros((tind.size, 2), dtype=np.float32)
    valid = w >= 0
    if not valid.any() or tind.size == 0:
        return out
    n_valid = int(valid.sum())
    norm = np.where(valid, w / w[valid].sum() * (n_valid if multiply else 1.0), 0.0)
    keep = max(1, int(np.ceil(n_valid / max(1.0, float(fraction)))))
    order = np.argsort(-np.where(valid, w, -np.inf), kind="stable")
    full_band = np.zeros(w.size, dtype=bool)
    full_band[order[:keep]] = True
    cut = np.where(full_band | ~valid, 0.0, float(transition) * float(r_rec))
    inside = (tind >= 0) & (tind < w.size)
    idx = np.clip(tind, 0, w.size - 1)
    out[:, 0] = np.where(inside, norm[idx], 0.0)
    out[:, 1] = np.where(inside, cut[idx], 0.0)
    return out


def parameter_statistics(rows):
    """Rows 0 / 1 of `<name>_stat.cistem`: np.mean / np.var (population) of every column, in float64, stored
    through the column types of the table like any other row (particle_cspt.py:1009-1016)."""
    out = np.zeros(2, dtype=rows.dtype)
    if rows.size == 0:
        return out
    for name in rows.dtype.names:
        col = rows[name].astype(np.float64)
        out[name][0] = np.mean(col)
        out[name][1] = np.var(col)
    return out


def statistics_from_moments(dtype, count, sums, sums_sq):
    """The same two rows from per-column (count, sum x, sum x^2) — what the ranks of a multi-GPU run add up
    (`dist.allreduce_parameter_statistics`); equal to `parameter_statistics` up to float64 rounding."""
    out = np.zeros(2, dtype=dtype)
    if count <= 0:
        return out
    for k, name in enumerate(np.dtype(dtype).names):
        mean = sums[k] / count
        out[name][0] = mean
        out[name][1] = max(sums_sq[k] / count - mean * mean, 0.0)
    return out
