"""The step between refine3d and reconstruct3d: score shaping (which projections enter the
reconstruction) and LogP -> occupancy for multi-class refinement.  SURVEY.md §8f rank 1.

Restates, vectorised over packed rows, what pyp does in
  * src/pyp/analysis/scores.py:300-761  ``shape_phase_residuals`` as called by
    ``call_shape_phase_residuals`` (:766-825): one angular and one defocus group (the defaults of
    ``reconstruct_agroups`` / ``reconstruct_dgroups``), scores (not phase residuals), a cutoff in (0, 1] or
    the automatic cutoff ``reconstruct_cutoff = 0`` (two-Gaussian fit of the score distribution,
    scores.py:438-465 with src/pyp/analysis/statistics.py:10-148);
  * src/pyp/analysis/occupancies.py:173-208 ``occupancy_extended`` (the LogP -> occupancy law).
Pinned against the reference's own outputs in tests/golden/shape_{spa,tomo,bimodal}_*.cistem.
"""
import numpy as np


def optimal_threshold(samples, random_state=0):
    """Score separating the two populations of a bimodal sample (statistics.py:10-148, criterion
    "optimal"): fit a two-component Gaussian mixture, evaluate both weighted components on 5000
    points between the sample extremes and take the point between the two means where they are
    closest.  Falls back to mean - 3 sigma of a single Gaussian when the components do not cross
    there or their sum exceeds both peaks (one population).  A constant sample returns 1.
    The reference leaves the k-means initialisation of the fit unseeded; `random_state` pins it here
    (the fixture was checked to be seed-independent, tests/golden/make_golden_bimodal.py).
    scikit-learn is the reference's own dependency for this step."""
    s = np.asarray(samples, dtype=np.float64).ravel()
    if s.size == 0 or np.var(s) == 0:
        return 1.0
    from sklearn.mixture import GaussianMixture

    trapezoid = getattr(np, "trapezoid", None) or np.trapz
    gm = GaussianMixture(n_components=2, covariance_type="full", tol=1e-6, reg_covar=1e-6, random_state=random_state).fit(s[:, None])
    mu, var, wt = gm.means_.ravel(), gm.covariances_.ravel(), gm.weights_.ravel()
    x = np.linspace(s.min(), s.max(), 5000)
    comps = []
    for k in range(2):
        g = np.exp(-((x - mu[k]) ** 2) / (2.0 * var[k]))
        comps.append(g / trapezoid(g, x) * wt[k])
    total = np.zeros_like(x, dtype=np.float32)  # the reference accumulates the sum in float32
    for c in comps:
        total += c
    lo, hi = sorted((int(np.argmin(np.abs(x - mu[1]))), int(np.argmin(np.abs(x - mu[0])))))
    if hi <= lo:
        crossing = False
        at = lo
    else:
        at = lo + int(np.argmin(np.abs(comps[0][lo:hi] - comps[1][lo:hi])))
        a, b = max(at - 1, 0), min(at + 1, x.size - 1)
        crossing = (comps[0][a] - comps[1][a]) * (comps[0][b] - comps[1][b]) <= 0
    one_population = total[at] > comps[0].max() and total[at] > comps[1].max()
    if not crossing or one_population:
        return float(s.mean() - 3.0 * np.sqrt(s.var() + 1e-6))  # single Gaussian, reg_covar as in the fit
    return float(x[at])


def shape_scores(rows, tilt_angle, cutoff, mindef=0.0, maxdef=100000.0, firstframe=0, lastframe=-1, mintilt=-90.0, maxtilt=90.0,
                 minazh=0.0, maxazh=180.0, minscore=0.0, maxscore=1.0, renumber=True):
    """Return a copy of `rows` with OCCUPANCY zeroed for the projections that must not enter the
    reconstruction.  `tilt_angle` = per-row tilt angle in degrees (all zero for single particle data,
    scores.py:340-372).  `renumber` rewrites POSITION_IN_STACK = 1..n as for `_used.cistem` files
    (scores.py:757-759)."""
    if not (0.0 <= cutoff <= 1.0):
        raise ValueError("reconstruct_cutoff must be 0 (automatic) or a fraction in (0, 1]; absolute counts > 1 "
                         "(scores.py:499-517) are not implemented")
    out = rows.copy()
    n = out.size
    tilt_angle = np.asarray(tilt_angle, dtype=np.float64).reshape(-1)
    if tilt_angle.size != n:
        raise ValueError("tilt_angle must have one entry per row")
    score = out["score"].astype(np.float64)
    occ = out["occupancy"].astype(np.float64)
    is_tomo = bool(np.any(np.abs(tilt_angle) > 0))
    if n == 0:
        return out

    def particle_means(sel):
        pind = out["pind"][sel]
        ids, inv = np.unique(pind, return_inverse=True)
        sums = np.bincount(inv, weights=score[sel], minlength=ids.size)
        cnt = np.bincount(inv, minlength=ids.size)
        return ids, sums / cnt

    # ---- per-cluster threshold (one cluster): scores.py:438-497
    threshold = np.nan
    if cutoff == 0:
        # automatic: 1.075 x the crossing of the two fitted score populations, if there are > 20 values
        prs = particle_means(np.abs(tilt_angle) <= 12)[1] if is_tomo else score
        thr = 1.075 * optimal_threshold(prs)
        if prs.size > 20:
            threshold = thr
    elif is_tomo:
        _, means = particle_means(np.abs(tilt_angle) <= 12)
        if means.size:
            threshold = np.sort(means)[int((means.size - 1) * (1 - cutoff))]
    else:
        threshold = np.sort(score)[int((n - 1) * (1 - cutoff))]
    lo = score.min() + minscore * (score.max() - score.min()) if minscore < 1 else minscore      # :519-523
    hi = score.max() - (1 - maxscore) * (score.max() - score.min()) if maxscore <= 1 else maxscore  # :525-530
    # ---- apply: scores.py:571-623
    if is_tomo and threshold > 0:
        near = np.abs(tilt_angle) < 10
        ids, means = particle_means(near)
        bad = ids[~(means >= threshold)] if cutoff != 1 else ids[:0]
        occ[np.isin(out["pind"], bad)] = 0
        occ[(score < lo) | (score > hi)] = 0
    else:
        occ[(score < threshold) | (score < lo) | (score > hi)] = 0
    # ---- windows: scores.py:646-690
    d1 = out["defocus_1"].astype(np.float64)
    occ[(d1 < mindef) | (d1 > maxdef)] = 0
    if maxazh < 180 or minazh > 0:
        az = np.mod(out["theta"].astype(np.float64), 180)  # column 2 of the table is THETA (scores.py:660)
        occ[(az < minazh) | (az > maxazh)] = 0
    if lastframe > -1:
        occ[(out["tind"] < firstframe) | (out["tind"] > lastframe)] = 0
    occ[(tilt_angle < mintilt) | (tilt_angle > maxtilt)] = 0
    out["occupancy"] = occ
    if renumber:
        out["position_in_stack"] = np.arange(1, n + 1)
    return out


def tilt_angles_of_rows(rows, tilts_by_film):
    """Per-row tilt angle from the `{film: {tind: angle}}` table pyp stores next to the merged
    parameter file (`<name>.json`, scores.py:336-361); films are the IMAGE_IS_ACTIVE column."""
    out = np.zeros(rows.size, dtype=np.float64)
    for k, r in enumerate(rows):
        film = tilts_by_film.get(str(int(r["image_is_active"])), {})
        out[k] = float(film.get(str(int(r["tind"])), film.get(int(r["tind"]), np.nan))) if film else np.nan
    return out


def class_occupancies(logp, sigma, class_average_occ):
    """LogP -> occupancy over K classes (occupancies.py:173-208).  logp, sigma: (K, n) arrays,
    class_average_occ: K mean occupancies of the previous iteration.  Returns (occ (K, n) in percent,
    sigma (n,)): occ_k = 100 a_k e^{-d_k} / sum_j a_j e^{-d_j}, d_k = max_j logp_j - logp_k, terms with
    d >= 10 dropped; sigma = sum_k sigma_k occ_k / 100."""
    logp = np.asarray(logp, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    a = np.asarray(class_average_occ, dtype=np.float64).reshape(-1, 1)
    if logp.ndim != 2 or logp.shape != sigma.shape or a.shape[0] != logp.shape[0]:
        raise ValueError("logp and sigma must be (classes, projections), one average occupancy per class")
    delta = logp.max(axis=0, keepdims=True) - logp
    keep = delta < 10
    pp = np.where(keep, np.exp(-delta) * a, 0.0)
    total = pp.sum(axis=0, keepdims=True)
    occ = np.where(keep, pp * 100.0 / total, 0.0)
    return occ, (sigma * occ / 100.0).sum(axis=0)
