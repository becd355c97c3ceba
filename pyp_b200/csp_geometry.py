"""Tilt/particle pose composition of constrained single-particle refinement (CSP).

The reference composes the pose of projection (particle p, tilt t) in
src/pyp/analysis/geometry/core.py:1081-1217 (`csp_euler_angles`) and stores the factors in the
`_extended.cistem` tables (cistem_star_file.py:247-248): particles hold (PPSI, PTHETA, PPHI,
PSHIFT_X/Y/Z) = minus the decoded particle matrix / the 3-D shifts, tilts hold (TILTANG,
TILTAXIS = -axis, TSHIFT_X/Y).  With pyp's left-handed matrix L (geometry/core.py:176-180) and
right-handed elementary rotations (its `vtk.rotation_matrix`):

    r_row            = L(-PPSI, -PTHETA, -PPHI) . Ry(TILTANG) . Rz(TILTAXIS)      (decoded by get_degrees_from_matrix)
    shift_row (x, y) = [ Rz(TILTAXIS) . Ry(TILTANG) . (-PSHIFT) ]_{x,y}

Both identities are pinned against the reference's own function in tests/golden/csp_euler.npy.
The kernels work in cisTEM's convention M(psi,theta,phi) = Rz(phi) Ry(theta) Rz(psi), related by
M(a) = D L(a) D, D = diag(-1, 1, 1), hence

    M_row = M(-PPSI, -PTHETA, -PPHI) . Ry(-TILTANG) . Rz(-TILTAXIS)

which is what `compose_pose` (and csp.cu / the oracle) evaluate.
"""
import numpy as np


def rz(a_deg):
    a = np.radians(a_deg)
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def ry(a_deg):
    a = np.radians(a_deg)
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def m_matrix(psi, theta, phi):
    """cisTEM / FREALIGN matrix Rz(phi) Ry(theta) Rz(psi) (SEMANTICS.md §2)."""
    return rz(phi) @ ry(theta) @ rz(psi)


def decode_m(m):
    """(psi, theta, phi) in degrees, psi/phi in [0, 360), of M = Rz(phi) Ry(theta) Rz(psi).
    Same branch structure as get_degrees_from_matrix (geometry/core.py:211-236)."""
    sth = np.hypot(m[2, 0], m[2, 1])
    if sth > 1e-7:
        theta = np.arctan2(sth, m[2, 2])
        psi = np.arctan2(m[2, 1], -m[2, 0])
        phi = np.arctan2(m[1, 2], m[0, 2])
    elif m[2, 2] > 0:
        theta, psi, phi = 0.0, 0.0, np.arctan2(m[1, 0], m[0, 0])
    else:
        theta, psi, phi = np.pi, 0.0, np.arctan2(-m[1, 0], -m[0, 0])
    out = np.degrees([psi, theta, phi])
    out[0] %= 360.0
    out[2] %= 360.0
    return out


def tilt_projector(angle, axis):
    """First two rows of Rz(TILTAXIS) Ry(TILTANG): 3-D offsets -> 2-D offsets on tilt image."""
    return (rz(axis) @ ry(angle))[:2]


def compose_pose(particle, tilt):
    """particle = (PPSI, PTHETA, PPHI), tilt = (TILTANG, TILTAXIS) -> row (psi, theta, phi)."""
    m = m_matrix(-particle[0], -particle[1], -particle[2]) @ ry(-tilt[0]) @ rz(-tilt[1])
    return decode_m(m)


def compose_shift(pshift, tilt):
    """2-D shift (same unit as pshift) that the 3-D particle shift induces on a tilt image."""
    return tilt_projector(tilt[0], tilt[1]) @ (-np.asarray(pshift, dtype=np.float64))


def defocus_offset_from_center(particle_xyz, tomo_center_xyz, tilt_angle_deg, specimen_z_offset, handedness=-1):
    """Height difference (tomogram voxels) between a particle and the tilt-series centre along the beam at
    one tilt — the per-particle, per-tilt defocus offset of tilt-series CTF (defocus += offset x pixel).
    Closed form of src/pyp/analysis/geometry/core.py:686-773 (`DefocusOffsetFromCenter`): the chain
    toRaw . T^-1 . Ry(tilt) . toOrigin applied to the particle and to the centre lifted by the specimen
    offset; the in-plane alignment T^-1 leaves z alone, so with d = particle - centre

        offset = -sin(tilt) d_x + cos(tilt) (d_z - specimen_z_offset),   times -1 for handedness = -1.

    Pinned against the reference's function in tests/golden/defocus_offset.npy."""
    if handedness not in (1, -1):
        raise ValueError("handedness must be 1 or -1")
    d = np.asarray(particle_xyz, dtype=np.float64) - np.asarray(tomo_center_xyz, dtype=np.float64)
    a = np.radians(tilt_angle_deg)
    off = -np.sin(a) * d[..., 0] + np.cos(a) * (d[..., 2] - specimen_z_offset)
    return -off if handedness == -1 else off
