"""Synthetic particle stacks for tests and the benchmark (SURVEY.md §8d "Synthetic inputs").

The phantom is a sum of isotropic 3-D Gaussians, so both the volume and every projection have
closed forms in real space: a blob centred at c (voxels from the box centre) appears in the
projection with pose matrix M = Rz(phi) Ry(theta) Rz(psi) at (M^T c)_{x,y}.  No FFT slicing or
interpolation is involved in making the data, which keeps it independent of both the CUDA engine
and the CPU oracle.  CTF and noise are applied with numpy (or torch on the GPU for large stacks —
data generation only, not the product path).
"""
import math

import numpy as np


def euler_matrix(psi, theta, phi):
    """FREALIGN/cisTEM ZYZ matrix, degrees (decode: src/pyp/analysis/geometry/core.py:222-247)."""
    ps, th, ph = (math.radians(v) for v in (psi, theta, phi))
    cps, sps, cth, sth, cph, sph = math.cos(ps), math.sin(ps), math.cos(th), math.sin(th), math.cos(ph), math.sin(ph)
    return np.array(
        [
            [cph * cth * cps - sph * sps, -cph * cth * sps - sph * cps, cph * sth],
            [sph * cth * cps + cph * sps, -sph * cth * sps + cph * cps, sph * sth],
            [-sth * cps, sth * sps, cth],
        ]
    )


class Phantom:
    """~`n_blobs` random Gaussians inside radius 0.35 n, seed 0."""

    def __init__(self, n, n_blobs=200, seed=0, radius_frac=0.35, sigma=2.0):
        rng = np.random.default_rng(seed)
        v = rng.normal(size=(n_blobs, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        r = radius_frac * n * rng.random(n_blobs) ** (1.0 / 3.0)
        self.n = n
        self.centres = v * r[:, None]  # (x, y, z) voxels from the box centre
        self.amps = 0.5 + rng.random(n_blobs)
        self.sigma = float(sigma)

    def volume(self):
        n, s = self.n, self.sigma
        ax = np.arange(n) - n // 2
        vol = np.zeros((n, n, n), dtype=np.float64)
        for (cx, cy, cz), a in zip(self.centres, self.amps):
            gx = np.exp(-((ax - cx) ** 2) / (2 * s * s))
            gy = np.exp(-((ax - cy) ** 2) / (2 * s * s))
            gz = np.exp(-((ax - cz) ** 2) / (2 * s * s))
            vol += a * gz[:, None, None] * gy[None, :, None] * gx[None, None, :]
        return vol.astype(np.float32)

    def project(self, psi, theta, phi, shift_x=0.0, shift_y=0.0):
        """Noise-free projection, shifts in pixels (particle displaced by +shift)."""
        n, s = self.n, self.sigma
        M = euler_matrix(psi, theta, phi)
        c2 = self.centres @ M  # rows = M^T c
        ax = np.arange(n) - n // 2
        ex = np.exp(-((ax[None, :] - (c2[:, 0:1] + shift_x)) ** 2) / (2 * s * s))
        ey = np.exp(-((ax[None, :] - (c2[:, 1:2] + shift_y)) ** 2) / (2 * s * s))
        w = self.amps * math.sqrt(2 * math.pi) * s
        return (ey * w[:, None]).T @ ex  # [y, x]


def ctf_2d(n, pixel, d1, d2, ast_deg, kv=300.0, cs_mm=2.7, ampl=0.07, phase_shift=0.0):
    """Full (n, n) CTF in numpy FFT order, float64 (oracle/SEMANTICS.md §CTF)."""
    v = kv * 1000.0
    lam = 12.2639 / math.sqrt(v + 0.97845e-6 * v * v)
    f = np.fft.fftfreq(n, d=pixel)
    fx, fy = np.meshgrid(f, f)  # fx varies along axis 1
    s2 = fx * fx + fy * fy
    ang = np.arctan2(fy, fx)
    df = 0.5 * (d1 + d2 + (d1 - d2) * np.cos(2 * (ang - math.radians(ast_deg))))
    chi = math.pi * lam * s2 * (df - 0.5 * lam * lam * s2 * cs_mm * 1e7) + phase_shift + math.atan(ampl / math.sqrt(1 - ampl * ampl))
    return -np.sin(chi)


def make_rows(n_part, pixel, seed=1, shift_px=5.0, defocus=(10000.0, 30000.0), dtype=None):
    """Random true poses / CTF per SURVEY.md §8d; shifts stored in Angstrom."""
    from ._lib import ROW_DTYPE

    rng = np.random.default_rng(seed)
    rows = np.zeros(n_part, dtype=dtype or ROW_DTYPE)
    rows["position_in_stack"] = np.arange(1, n_part + 1)
    rows["psi"] = rng.uniform(0, 360, n_part)
    rows["theta"] = np.degrees(np.arccos(rng.uniform(-1, 1, n_part)))
    rows["phi"] = rng.uniform(0, 360, n_part)
    rows["x_shift"] = rng.uniform(-shift_px, shift_px, n_part) * pixel
    rows["y_shift"] = rng.uniform(-shift_px, shift_px, n_part) * pixel
    d = rng.uniform(defocus[0], defocus[1], n_part)
    a = rng.normal(0, 500.0, n_part)
    rows["defocus_1"] = d + a / 2
    rows["defocus_2"] = d - a / 2
    rows["defocus_angle"] = rng.uniform(0, 180, n_part)
    rows["occupancy"] = 100.0
    rows["sigma"] = 0.5
    rows["score"] = 0.5
    rows["pixel_size"] = pixel
    rows["voltage_kv"] = 300.0
    rows["cs_mm"] = 2.7
    rows["amplitude_contrast"] = 0.07
    rows["image_is_active"] = 1
    rows["pind"] = np.arange(n_part)
    return rows


def make_stack(phantom, rows, snr=0.05, seed=2, apply_ctf=True):
    """Noisy CTF-modulated projections (n_part, n, n) float32."""
    n = phantom.n
    rng = np.random.default_rng(seed)
    out = np.zeros((rows.size, n, n), dtype=np.float32)
    for k, r in enumerate(rows):
        px = float(r["pixel_size"])
        img = phantom.project(r["psi"], r["theta"], r["phi"], r["x_shift"] / px, r["y_shift"] / px)
        if apply_ctf:
            c = ctf_2d(n, px, r["defocus_1"], r["defocus_2"], r["defocus_angle"], r["voltage_kv"], r["cs_mm"], r["amplitude_contrast"], r["phase_shift"])
            img = np.fft.ifft2(np.fft.fft2(img) * c).real
        if snr is not None and snr > 0:
            sig = img.std()
            img = img + rng.normal(0, sig / math.sqrt(snr), img.shape)
        out[k] = img
    return out


def perturb_rows(rows, ang_sigma=2.0, shift_sigma_px=1.0, seed=3):
    """Input table for refinement: true poses jittered by N(0, 2 deg) / N(0, 1 px) (§8d)."""
    rng = np.random.default_rng(seed)
    out = rows.copy()
    for k in ("psi", "theta", "phi"):
        out[k] = rows[k] + rng.normal(0, ang_sigma, rows.size)
    out["x_shift"] = rows["x_shift"] + rng.normal(0, shift_sigma_px, rows.size) * rows["pixel_size"]
    out["y_shift"] = rows["y_shift"] + rng.normal(0, shift_sigma_px, rows.size) * rows["pixel_size"]
    return out
