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


def make_tilt_series(n_particles, pixel, tilt_angles=None, tilt_axis=0.0, seed=1, shift_a=6.0, defocus=25000.0,
                     extent_px=300.0, thickness_px=60.0, handedness=1.0, dtype=None):
    """Synthetic sub-tomogram tilt series (BASELINE configs[2]): every particle is seen at every
    tilt; the projection rows are composed from the extended tables exactly as the reference does
    (pyp_b200/csp_geometry.py, pinned to geometry/core.py:1081-1217) and carry a per-tilt,
    per-particle defocus following inout/metadata/core.py:2831-2865.

    Returns (rows, particles, tilts); rows are particle-major (all tilts of particle 0 first),
    POSITION_IN_STACK 1..n, TIND = acquisition (scan) order index of the tilt."""
    from . import csp_geometry as G
    from .formats.cistem import PARTICLE_DTYPE, TILT_DTYPE
    from ._lib import ROW_DTYPE

    rng = np.random.default_rng(seed)
    if tilt_angles is None:
        tilt_angles = np.arange(-60.0, 60.1, 3.0)
    tilt_angles = np.asarray(tilt_angles, dtype=np.float64)
    nt = tilt_angles.size
    # dose-symmetric acquisition order: 0, +3, -3, +6, -6 ...  (scan-order index = TIND)
    order = np.argsort(np.abs(tilt_angles) + 1e-3 * (tilt_angles < 0), kind="stable")
    tind_of = np.empty(nt, dtype=int)
    tind_of[order] = np.arange(nt)
    tilts = np.zeros(nt, dtype=TILT_DTYPE)
    for k in range(nt):
        tilts[tind_of[k]] = (tind_of[k], 0, 0.0, 0.0, tilt_angles[k], tilt_axis)
    particles = np.zeros(n_particles, dtype=PARTICLE_DTYPE)
    particles["pind"] = np.arange(n_particles)
    particles["psi"] = rng.uniform(0, 360, n_particles)
    particles["theta"] = np.degrees(np.arccos(rng.uniform(-1, 1, n_particles)))
    particles["phi"] = rng.uniform(0, 360, n_particles)
    for k in ("shift_x", "shift_y", "shift_z"):
        particles[k] = rng.uniform(-shift_a, shift_a, n_particles)
    particles["x_position_3d"] = rng.uniform(-extent_px, extent_px, n_particles) + 2000.0
    particles["y_position_3d"] = rng.uniform(-extent_px, extent_px, n_particles) + 2000.0
    particles["z_position_3d"] = rng.uniform(-thickness_px, thickness_px, n_particles) + 150.0
    particles["occ"] = 100.0
    centre = np.array([particles["x_position_3d"].mean(), particles["y_position_3d"].mean(), particles["z_position_3d"].mean()])
    rows = np.zeros(n_particles * nt, dtype=dtype or ROW_DTYPE)
    k = 0
    for p in particles:
        pos = (np.array([p["x_position_3d"], p["y_position_3d"], p["z_position_3d"]]) - centre) * pixel
        for t in tilts:
            ang = G.compose_pose((p["psi"], p["theta"], p["phi"]), (t["angle"], t["axis"]))
            sh = G.compose_shift((p["shift_x"], p["shift_y"], p["shift_z"]), (t["angle"], t["axis"]))
            r = rows[k]
            r["position_in_stack"] = k + 1
            r["psi"], r["theta"], r["phi"] = ang
            r["x_shift"], r["y_shift"] = sh + (t["shift_x"], t["shift_y"])
            # height of the particle above the tilt axis plane changes the defocus
            # (inout/metadata/core.py:2857-2865: -z cos(t) + x sin(t) terms)
            a = math.radians(handedness * t["angle"])
            dz = -pos[2] * math.cos(a) + pos[0] * math.sin(a)
            r["defocus_1"] = defocus + dz + 150.0
            r["defocus_2"] = defocus + dz - 150.0
            r["defocus_angle"] = 35.0
            r["occupancy"], r["sigma"], r["score"] = 100.0, 1.0, 0.5
            r["pixel_size"], r["voltage_kv"], r["cs_mm"], r["amplitude_contrast"] = pixel, 300.0, 2.7, 0.07
            r["image_is_active"] = 0
            r["pind"], r["tind"], r["rind"], r["find"] = p["pind"], t["tind"], 0, 0
            r["imind"] = t["tind"]
            r["original_x"], r["original_y"] = p["x_position_3d"], p["y_position_3d"]
            k += 1
    return rows, particles, tilts


def perturb_particles(particles, ang_sigma=2.0, shift_sigma_a=1.5, seed=5):
    rng = np.random.default_rng(seed)
    out = particles.copy()
    for k in ("psi", "theta", "phi"):
        out[k] = particles[k] + rng.normal(0, ang_sigma, particles.size)
    for k in ("shift_x", "shift_y", "shift_z"):
        out[k] = particles[k] + rng.normal(0, shift_sigma_a, particles.size)
    return out


def rows_from_tables(rows, particles0, tilts0, particles, tilts):
    """Re-compose every projection row after its particle / tilt parameters changed (float64
    restatement of the CSP pose model, SEMANTICS.md §11) — used to build perturbed inputs."""
    from . import csp_geometry as G

    out = rows.copy()
    pi = {int(p["pind"]): k for k, p in enumerate(particles)}
    ti = {(int(t["tind"]), int(t["rind"])): k for k, t in enumerate(tilts)}
    centre = np.array([particles["x_position_3d"].mean(), particles["y_position_3d"].mean(), particles["z_position_3d"].mean()], dtype=np.float64)
    for r in out:
        p, p0 = particles[pi[int(r["pind"])]], particles0[pi[int(r["pind"])]]
        t, t0 = tilts[ti[(int(r["tind"]), int(r["rind"]))]], tilts0[ti[(int(r["tind"]), int(r["rind"]))]]
        X = (np.array([p["x_position_3d"], p["y_position_3d"], p["z_position_3d"]], dtype=np.float64) - centre) * float(r["pixel_size"])
        v = X - np.array([p["shift_x"], p["shift_y"], p["shift_z"]], dtype=np.float64)
        v0 = X - np.array([p0["shift_x"], p0["shift_y"], p0["shift_z"]], dtype=np.float64)
        s = G.tilt_projector(t["angle"], t["axis"]) @ v - G.tilt_projector(t0["angle"], t0["axis"]) @ v0
        r["psi"], r["theta"], r["phi"] = G.compose_pose((p["psi"], p["theta"], p["phi"]), (t["angle"], t["axis"]))
        r["x_shift"] += s[0] + t["shift_x"] - t0["shift_x"]
        r["y_shift"] += s[1] + t["shift_y"] - t0["shift_y"]
    return out
