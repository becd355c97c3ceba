"""Beam-tilt estimation of refine_ctf (answer 23 "estimate beam tilt";
src/pyp/refine/frealign/frealign.py:3995-4041).

The GPU accumulates S(g) = sum over particles of G * conj(CTF * slice) on the scoring band
(`cspb_refine_phase_sum`, the same gather as one score evaluation per particle).  A tilted beam adds
the odd phase error (axial coma)

    phi(g) = 2 pi Cs lambda^2 |g|^2 (g . b)            b = beam tilt (rad), g in 1/Angstrom,

and a residual common displacement t (Angstrom) adds 2 pi (g . t).  Both are linear in the four
unknowns, which are fitted by least squares on arg S, weighted by |S| — first on the low-phase part
of the band, then re-fitted on the residual unwrapped around the model (oracle/SEMANTICS.md §12).
The three diagnostic images cisTEM writes (phase difference, fitted beam-tilt phase, their
difference) are returned as full n x n images in FFT-shifted order.
"""
import math

import numpy as np


def wavelength(kv):
    v = kv * 1000.0
    return 12.2639 / math.sqrt(v + 0.97845e-6 * v * v)


def _grid(n, pixel):
    jj, ii = np.meshgrid(np.arange(n), np.arange(n // 2 + 1), indexing="ij")
    j = np.where(jj >= n // 2, jj - n, jj)
    gx = ii / (n * pixel)
    gy = j / (n * pixel)
    return gx, gy


def tilt_phase(n, pixel, kv, cs_mm, beam_tilt_mrad, shift_a=(0.0, 0.0)):
    """Model phase (radians) on the n x (n/2+1) half-plane grid."""
    gx, gy = _grid(n, pixel)
    lam, cs = wavelength(kv), cs_mm * 1.0e7
    g2 = gx * gx + gy * gy
    bx, by = beam_tilt_mrad[0] * 1e-3, beam_tilt_mrad[1] * 1e-3
    return 2 * math.pi * cs * lam * lam * g2 * (gx * bx + gy * by) + 2 * math.pi * (gx * shift_a[0] + gy * shift_a[1])


def fit(S, pixel, kv, cs_mm, passes=3):
    """Least-squares beam tilt (mrad) and common shift (Angstrom) from the phase sum S (n x n/2+1 complex)."""
    S = np.asarray(S)
    n = S.shape[0]
    gx, gy = _grid(n, pixel)
    lam, cs = wavelength(kv), cs_mm * 1.0e7
    g2 = gx * gx + gy * gy
    k = 2 * math.pi * cs * lam * lam
    w = np.abs(S).astype(np.float64)
    m = w > 0
    # the i = 0 column holds both members of each Friedel pair: keep j >= 0 there
    m &= ~((gx == 0) & (gy < 0))
    A = np.stack([k * g2 * gx, k * g2 * gy, 2 * math.pi * gx, 2 * math.pi * gy], axis=-1)[m]
    y = np.angle(S[m]).astype(np.float64)
    ww = w[m]
    x = np.zeros(4)
    if ww.size < 8:
        return dict(beam_tilt_x=0.0, beam_tilt_y=0.0, shift_x=0.0, shift_y=0.0, rms=0.0, phase=np.zeros((n, n), np.float32),
                    model=np.zeros((n, n), np.float32), difference=np.zeros((n, n), np.float32))
    for _ in range(max(1, passes)):
        r = y - A @ x
        r = (r + math.pi) % (2 * math.pi) - math.pi  # unwrap around the current model
        Aw = A * ww[:, None]
        dx, *_ = np.linalg.lstsq(A.T @ Aw, Aw.T @ r, rcond=None)
        x = x + dx
    r = (y - A @ x + math.pi) % (2 * math.pi) - math.pi
    rms = float(np.sqrt((ww * r * r).sum() / ww.sum()))
    bt = (x[0] * 1e3, x[1] * 1e3)
    model_h = np.where(w > 0, tilt_phase(n, pixel, kv, cs_mm, bt), 0.0)
    phase_h = np.where(w > 0, np.angle(S), 0.0)
    return dict(beam_tilt_x=float(bt[0]), beam_tilt_y=float(bt[1]), shift_x=float(x[2]), shift_y=float(x[3]), rms=rms,
                phase=_full(phase_h), model=_full(model_h), difference=_full(np.where(w > 0, phase_h - model_h, 0.0)))


def _full(half):
    """Odd (phase-like) extension of an n x (n/2+1) half-plane image to n x n, origin at (n/2, n/2)."""
    n = half.shape[0]
    full = np.zeros((n, n), dtype=np.float32)
    full[:, : n // 2 + 1] = half
    i = np.arange(n // 2 + 1, n)
    jm = (-np.arange(n)) % n
    full[:, i] = -half[jm][:, n - i]
    return np.fft.fftshift(full).astype(np.float32)


def apply_to_stack(stack, pixel, kv, cs_mm, beam_tilt_mrad):
    """Synthetic data: multiply the transform of every image by exp(i phi) (tests, KAT)."""
    n = stack.shape[-1]
    ph = tilt_phase(n, pixel, kv, cs_mm, beam_tilt_mrad)
    F = np.fft.rfft2(stack.astype(np.float64)) * np.exp(1j * ph)
    return np.fft.irfft2(F, s=(n, n)).astype(np.float32)
