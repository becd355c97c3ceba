"""GPU-side synthetic stack generator for the benchmark (data generation only — torch is used
here as plumbing to fill HBM with particles; nothing in this file is on the product path).

Same analytic model as :mod:`pyp_b200.synth`: Gaussian-blob phantom, closed-form projections,
CTF from the float formula, white noise at a given SNR.
"""
import math

import numpy as np
import torch

from .symmetry import symmetry_matrices


def symmetric_phantom(n, symbol="O", n_base=9, seed=0, radius_frac=0.35, sigma=2.0):
    """Blob centres/amplitudes replicated by the point group (apoferritin-like for 'O')."""
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n_base, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    r = radius_frac * n * (0.3 + 0.7 * rng.random(n_base))
    base = v * r[:, None]
    amps = 0.5 + rng.random(n_base)
    mats = symmetry_matrices(symbol).astype(np.float64)
    centres = np.concatenate([base @ m.T for m in mats], axis=0)
    return centres, np.tile(amps, mats.shape[0]), float(sigma)


def volume(n, centres, amps, sigma, device):
    ax = torch.arange(n, device=device, dtype=torch.float32) - n // 2
    c = torch.as_tensor(centres, device=device, dtype=torch.float32)
    a = torch.as_tensor(amps, device=device, dtype=torch.float32)
    gx = torch.exp(-((ax[None, :] - c[:, 0:1]) ** 2) / (2 * sigma * sigma))
    gy = torch.exp(-((ax[None, :] - c[:, 1:2]) ** 2) / (2 * sigma * sigma))
    gz = torch.exp(-((ax[None, :] - c[:, 2:3]) ** 2) / (2 * sigma * sigma))
    return torch.einsum("g,gz,gy,gx->zyx", a, gz, gy, gx).contiguous()


def _euler_batch(psi, theta, phi):
    ps, th, ph = (torch.deg2rad(t) for t in (psi, theta, phi))
    cps, sps, cth, sth, cph, sph = ps.cos(), ps.sin(), th.cos(), th.sin(), ph.cos(), ph.sin()
    m = torch.stack(
        [
            cph * cth * cps - sph * sps, -cph * cth * sps - sph * cps, cph * sth,
            sph * cth * cps + cph * sps, -sph * cth * sps + cph * cps, sph * sth,
            -sth * cps, sth * sps, cth,
        ],
        dim=-1,
    )
    return m.reshape(-1, 3, 3)


@torch.no_grad()
def make_stack(n, centres, amps, sigma, rows, snr=0.05, seed=2, device="cuda", chunk=2048, out=None):
    """rows: numpy structured array (ROW_DTYPE).  Returns float32 CUDA tensor (n_part, n, n)."""
    n_part = rows.size
    dev = torch.device(device)
    if out is None:
        out = torch.empty((n_part, n, n), device=dev, dtype=torch.float32)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    c = torch.as_tensor(centres, device=dev, dtype=torch.float32)
    a = torch.as_tensor(amps, device=dev, dtype=torch.float32) * math.sqrt(2 * math.pi) * sigma
    ax = torch.arange(n, device=dev, dtype=torch.float32) - n // 2
    f = torch.fft.fftfreq(n, device=dev)
    fy, fx = torch.meshgrid(f, f, indexing="ij")

    def col(name, s, e):
        return torch.as_tensor(np.ascontiguousarray(rows[name][s:e]).astype(np.float32), device=dev)

    for s in range(0, n_part, chunk):
        e = min(n_part, s + chunk)
        px = col("pixel_size", s, e)
        M = _euler_batch(col("psi", s, e), col("theta", s, e), col("phi", s, e))
        c2 = torch.einsum("gk,bkj->bgj", c, M)  # rows of (M^T c)
        ux = c2[:, :, 0] + (col("x_shift", s, e) / px)[:, None]
        uy = c2[:, :, 1] + (col("y_shift", s, e) / px)[:, None]
        ex = torch.exp(-((ax[None, None, :] - ux[:, :, None]) ** 2) / (2 * sigma * sigma))
        ey = torch.exp(-((ax[None, None, :] - uy[:, :, None]) ** 2) / (2 * sigma * sigma)) * a[None, :, None]
        img = torch.bmm(ey.transpose(1, 2), ex)  # [b, y, x]
        # CTF (oracle/SEMANTICS.md §CTF), frequencies in 1/Angstrom
        v = col("voltage_kv", s, e) * 1000.0
        lam = 12.2639 / torch.sqrt(v + 0.97845e-6 * v * v)
        s2 = (fx * fx + fy * fy)[None] / (px * px)[:, None, None]
        ang = torch.atan2(fy, fx)[None]
        d1, d2 = col("defocus_1", s, e)[:, None, None], col("defocus_2", s, e)[:, None, None]
        ast = torch.deg2rad(col("defocus_angle", s, e))[:, None, None]
        df = 0.5 * (d1 + d2 + (d1 - d2) * torch.cos(2 * (ang - ast)))
        w = col("amplitude_contrast", s, e)
        lam_ = lam[:, None, None]
        chi = math.pi * lam_ * s2 * (df - 0.5 * lam_ * lam_ * s2 * (col("cs_mm", s, e) * 1e7)[:, None, None])
        chi = chi + col("phase_shift", s, e)[:, None, None] + torch.atan(w / torch.sqrt(1 - w * w))[:, None, None]
        img = torch.fft.ifft2(torch.fft.fft2(img) * (-torch.sin(chi))).real
        if snr and snr > 0:
            sig = img.std(dim=(1, 2), keepdim=True)
            img = img + torch.randn(img.shape, device=dev, generator=gen) * (sig / math.sqrt(snr))
        out[s:e] = img
    return out
