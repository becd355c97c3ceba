"""Likelihood blurring of reconstruct3d (answer 34 "apply likelihood blurring", `reconstruct_lblur`,
src/pyp/refine/frealign/frealign.py:1772,1817): every projection enters the reconstruction at a fan of
in-plane rotations around its refined psi, each weighted by its likelihood relative to the best one.

Only the yes/no answer reaches the binary; the fan is FREALIGN's legacy default (`lblur_start -10`,
`lblur_step 1`, `lblur_nrot 21`, `lblur_range 20`, frealign.py:766-770).  Ours (oracle/SEMANTICS.md §8b): the
likelihood ratio comes from the score with the LOGP law of §6, LogP_k - LogP_max =
-(n_s / 2) ln((1 - cc_k^2) / (1 - cc_max^2)), cc = SCORE / 100; members more than `range` below the best are
dropped; weights are normalised to 1 per projection and multiply its occupancy.  Built from the existing
device calls (`cspb_refine_score_poses`, `cspb_recon_insert`): host-side composition, no new kernel.
"""
import numpy as np

LBLUR_START, LBLUR_STEP, LBLUR_NROT, LBLUR_RANGE = -10.0, 1.0, 21, 20.0


def offsets(start=LBLUR_START, step=LBLUR_STEP, nrot=LBLUR_NROT):
    return start + step * np.arange(int(nrot), dtype=np.float64)


def weights(scores, n_samples, logp_range=LBLUR_RANGE):
    """scores: (n, K) SCORE values (100 x CC) of the K fan members -> (n, K) weights, rows summing to 1."""
    cc = np.clip(np.asarray(scores, dtype=np.float64) / 100.0, -0.999999, 0.999999)
    logp = -0.5 * float(n_samples) * np.log1p(-cc * cc)
    logp = np.where(cc > 0, logp, -np.inf)                 # an anti-correlated member never counts
    best = logp.max(axis=1, keepdims=True)
    best = np.where(np.isfinite(best), best, 0.0)
    d = best - logp
    w = np.where(d <= logp_range, np.exp(-np.minimum(d, 700.0)), 0.0)
    total = w.sum(axis=1, keepdims=True)
    k0 = int(np.argmin(np.abs(offsets(nrot=w.shape[1])))) if w.shape[1] == LBLUR_NROT else w.shape[1] // 2
    w[(total == 0).ravel(), k0] = 1.0                      # nothing correlates: keep the refined pose alone
    return w / w.sum(axis=1, keepdims=True)


def fan_poses(rows, deltas):
    """(n*K, 6) poses {psi + delta_k, theta, phi, x, y, 0} image-major, and the matching image index."""
    n, K = rows.size, len(deltas)
    poses = np.zeros((n, K, 6), dtype=np.float32)
    poses[:, :, 0] = np.mod(rows["psi"].astype(np.float64)[:, None] + np.asarray(deltas)[None, :], 360.0)
    for c, name in ((1, "theta"), (2, "phi"), (3, "x_shift"), (4, "y_shift")):
        poses[:, :, c] = rows[name][:, None]
    return poses.reshape(-1, 6), np.repeat(np.arange(n, dtype=np.int32), K)


def insert_blurred(eng, images, rows, n_samples, deltas=None, logp_range=LBLUR_RANGE, weight_cut=None):
    """Score the fan on the scorer side of `eng` (configured, reference set) and insert every member with its
    weight.  Returns the (n, K) weights."""
    deltas = offsets() if deltas is None else np.asarray(deltas, dtype=np.float64)
    eng.load_images(images)
    poses, idx = fan_poses(rows, deltas)
    sc = eng.score_poses(rows, idx, poses).reshape(rows.size, len(deltas))
    w = weights(sc, n_samples, logp_range)
    for k, d in enumerate(deltas):
        if not (w[:, k] > 0).any():
            continue
        member = rows.copy()
        member["psi"] = np.mod(rows["psi"].astype(np.float64) + d, 360.0)
        member["occupancy"] = rows["occupancy"] * w[:, k]
        if weight_cut is None:
            eng.recon_insert(images, member)
        else:
            eng.recon_insert(images, member, weight_cut)
    return w
