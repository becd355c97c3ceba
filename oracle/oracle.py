"""ctypes face of oracle/liboracle.so (cspb_oracle.c) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  "Parity unpinned": see cspb_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

ROW_DTYPE = np.dtype(
    [
        ("position_in_stack", "<u4"),
        ("psi", "<f4"), ("theta", "<f4"), ("phi", "<f4"),
        ("x_shift", "<f4"), ("y_shift", "<f4"),
        ("defocus_1", "<f4"), ("defocus_2", "<f4"), ("defocus_angle", "<f4"), ("phase_shift", "<f4"),
        ("image_is_active", "<i4"),
        ("occupancy", "<f4"), ("logp", "<f4"), ("sigma", "<f4"), ("score", "<f4"),
        ("pixel_size", "<f4"), ("voltage_kv", "<f4"), ("cs_mm", "<f4"), ("amplitude_contrast", "<f4"),
        ("beam_tilt_x", "<f4"), ("beam_tilt_y", "<f4"), ("image_shift_x", "<f4"), ("image_shift_y", "<f4"),
        ("original_x", "<f4"), ("original_y", "<f4"),
        ("imind", "<i4"), ("pind", "<i4"), ("tind", "<i4"), ("rind", "<i4"), ("find", "<i4"),
        ("fshift_x", "<f4"), ("fshift_y", "<f4"),
    ]
)


class RefineCfg(C.Structure):
    _fields_ = [
        ("box", C.c_int32), ("pad", C.c_int32),
        ("pixel_size", C.c_float), ("mask_radius", C.c_float), ("low_res_limit", C.c_float),
        ("high_res_limit", C.c_float), ("signed_cc_limit", C.c_float), ("defocus_step", C.c_float),
        ("refine_psi", C.c_int32), ("refine_theta", C.c_int32), ("refine_phi", C.c_int32),
        ("refine_x", C.c_int32), ("refine_y", C.c_int32), ("refine_defocus", C.c_int32),
        ("apply_mask", C.c_int32), ("normalize", C.c_int32), ("invert_contrast", C.c_int32),
        ("whiten", C.c_int32), ("local_iterations", C.c_int32),
        ("search_high_res", C.c_float), ("search_range_x", C.c_float), ("search_range_y", C.c_float),
        ("best_matches", C.c_int32), ("global_search", C.c_int32),
        ("use_priors", C.c_int32), ("prior_mean_x", C.c_float), ("prior_mean_y", C.c_float),
        ("prior_var_x", C.c_float), ("prior_var_y", C.c_float),
        ("focus_x", C.c_float), ("focus_y", C.c_float), ("focus_z", C.c_float), ("focus_radius", C.c_float),
        ("optimizer", C.c_int32),
    ]


class ReconCfg(C.Structure):
    _fields_ = [
        ("box", C.c_int32), ("pad", C.c_int32),
        ("pixel_size", C.c_float), ("mask_radius", C.c_float), ("resolution_limit", C.c_float),
        ("score_bfactor", C.c_float), ("score_weighting", C.c_int32), ("score_threshold", C.c_float),
        ("normalize", C.c_int32), ("invert_contrast", C.c_int32), ("per_particle_split", C.c_int32),
        ("average_score", C.c_float),
    ]


PARTICLE_DTYPE = np.dtype([("pind", "<i4"), ("shift_x", "<f4"), ("shift_y", "<f4"), ("shift_z", "<f4"), ("psi", "<f4"), ("theta", "<f4"),
                           ("phi", "<f4"), ("x_position_3d", "<f4"), ("y_position_3d", "<f4"), ("z_position_3d", "<f4"), ("score", "<f4"), ("occ", "<f4")])
TILT_DTYPE = np.dtype([("tind", "<i4"), ("rind", "<i4"), ("shift_x", "<f4"), ("shift_y", "<f4"), ("angle", "<f4"), ("axis", "<f4")])


class CspCfg(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("window_min", C.c_int32), ("window_max", C.c_int32), ("iterations", C.c_int32),
        ("random_evals", C.c_int32), ("grid_search", C.c_int32),
        ("angle_step", C.c_float), ("shift_step", C.c_float),
        ("tol_particle_psi", C.c_float), ("tol_particle_theta", C.c_float), ("tol_particle_phi", C.c_float),
        ("tol_particle_shift", C.c_float),
        ("tol_tilt_angle", C.c_float), ("tol_tilt_axis", C.c_float), ("tol_tilt_shift", C.c_float), ("tol_defocus", C.c_float),
        ("seed", C.c_uint32), ("min_projections", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


def csp_cfg_from(gpu_cfg) -> CspCfg:
    o = CspCfg()
    for name, _ in CspCfg._fields_:
        if name != "reserved":
            setattr(o, name, getattr(gpu_cfg, name))
    return o


def refine_cfg_from(gpu_cfg) -> RefineCfg:
    """Copy the shared fields of a pyp_b200 RefineCfg (or any object with those attributes)."""
    o = RefineCfg()
    for name, _ in RefineCfg._fields_:
        setattr(o, name, getattr(gpu_cfg, name, 0))
    return o


def recon_cfg_from(gpu_cfg) -> ReconCfg:
    o = ReconCfg()
    for name, _ in ReconCfg._fields_:
        setattr(o, name, getattr(gpu_cfg, name))
    return o


def build():
    """(Re)build liboracle.so with the committed Makefile."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        sig = {
            "orc_fft2_r2c": (None, [vp, i, vp]),
            "orc_fft2_c2r": (None, [vp, i, vp]),
            "orc_euler_matrix": (None, [f, f, f, vp]),
            "orc_ctf_image": (None, [vp, i, vp]),
            "orc_band_limits": (f, [C.POINTER(RefineCfg), C.POINTER(f), C.POINTER(f)]),
            "orc_band_count": (i, [C.POINTER(RefineCfg)]),
            "orc_ref_create": (vp, [vp, i, i]),
            "orc_ref_free": (None, [vp]),
            "orc_project": (None, [vp, f, f, f, f, vp]),
            "orc_noise_curve": (None, [vp, i, C.POINTER(RefineCfg), vp]),
            "orc_prepare_image": (None, [vp, C.POINTER(RefineCfg), vp, vp, vp]),
            "orc_prepare_image_fast": (None, [vp, C.POINTER(RefineCfg), vp, vp, vp]),
            "orc_score": (f, [vp, vp, vp, vp, C.POINTER(RefineCfg), vp]),
            "orc_score_grad": (f, [vp, vp, vp, vp, C.POINTER(RefineCfg), vp, vp, vp, vp]),
            "orc_score_grad_cut": (f, [vp, vp, vp, vp, C.POINTER(RefineCfg), vp, vp, vp, vp, i]),
            "orc_refine_local_fast": (C.c_longlong, [vp, vp, vp, i, C.POINTER(RefineCfg)]),
            "orc_recon_insert_fast": (None, [vp, vp, vp, i, vp, i, i]),
            "orc_recon_finish_fast": (None, [vp]),
            "orc_recon_discard_fast": (None, [vp]),
            "orc_refine_local": (C.c_longlong, [vp, vp, vp, i, C.POINTER(RefineCfg)]),
            "orc_normalize": (None, [vp, i, f, i, i, vp]),
            "orc_phase_sum": (None, [vp, vp, vp, i, C.POINTER(RefineCfg), vp]),
            "orc_focus_center": (None, [C.POINTER(RefineCfg), vp, vp, vp]),
            "orc_focus_logp": (f, [vp, vp, vp, vp, C.POINTER(RefineCfg), vp]),
            "orc_global_search": (C.c_longlong, [vp, vp, vp, i, C.POINTER(RefineCfg), vp, i]),
            "orc_csp_compose": (None, [vp, vp, vp, vp, vp, f, f, f, vp]),
            "orc_csp_run": (C.c_longlong, [vp, vp, vp, i, vp, i, vp, i, C.POINTER(RefineCfg), C.POINTER(CspCfg), i, i]),
            "orc_recon_create": (vp, [C.POINTER(ReconCfg)]),
            "orc_recon_free": (None, [vp]),
            "orc_recon_insert": (None, [vp, vp, vp, i, vp, i]),
            "orc_recon_insert_weighted": (None, [vp, vp, vp, i, vp, i, vp]),
            "orc_recon_get_dump": (None, [vp, i, vp]),
            "orc_recon_finalize": (None, [vp, f, f, vp, vp, vp, vp]),
            "orc_fsc": (None, [vp, vp, i, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def fft2_r2c(img):
    img = _f32(img)
    n = img.shape[0]
    out = np.zeros((n, n // 2 + 1), dtype=np.complex64)
    lib().orc_fft2_r2c(_p(img), n, _p(out))
    return out


def fft2_c2r(spec):
    spec = np.ascontiguousarray(spec, dtype=np.complex64)
    n = spec.shape[0]
    out = np.zeros((n, n), dtype=np.float32)
    lib().orc_fft2_c2r(_p(spec), n, _p(out))
    return out


def euler_matrix(psi, theta, phi):
    out = np.zeros(9, dtype=np.float32)
    lib().orc_euler_matrix(psi, theta, phi, _p(out))
    return out.reshape(3, 3)


def ctf_image(row, n):
    row = np.ascontiguousarray(row, dtype=ROW_DTYPE).reshape(1)
    out = np.zeros((n, n // 2 + 1), dtype=np.float32)
    lib().orc_ctf_image(_p(row), n, _p(out))
    return out


def band_limits(cfg):
    lo, hi = C.c_float(), C.c_float()
    lib().orc_band_limits(C.byref(cfg), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def band_count(cfg):
    return lib().orc_band_count(C.byref(cfg))


class Reference:
    def __init__(self, vol, pad=1):
        vol = _f32(vol)
        self.n = vol.shape[0]
        self.pad = pad
        self._h = lib().orc_ref_create(_p(vol), self.n, pad)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ref_free(self._h)
            self._h = None

    def project(self, psi, theta, phi, r_hi):
        out = np.zeros((self.n, self.n // 2 + 1), dtype=np.complex64)
        lib().orc_project(self._h, psi, theta, phi, r_hi, _p(out))
        return out


def noise_curve(imgs, cfg):
    imgs = _f32(imgs)
    out = np.zeros(cfg.box + 1, dtype=np.float32)
    lib().orc_noise_curve(_p(imgs), imgs.shape[0], C.byref(cfg), _p(out))
    return out


def prepare_images(imgs, cfg, curve=None, ring_weights=None, fast=False):
    """fast=True: the optimised CPU leg's single-precision transforms (cspb_oracle_fast.c), results equal to ~1e-6."""
    imgs = _f32(imgs)
    n = cfg.box
    out = np.zeros((imgs.shape[0], n, n // 2 + 1), dtype=np.complex64)
    curve = None if curve is None else _f32(curve)
    rw = None if ring_weights is None else _f32(ring_weights)
    fn = lib().orc_prepare_image_fast if fast else lib().orc_prepare_image
    for k in range(imgs.shape[0]):
        fn(_p(imgs[k]), C.byref(cfg), _p(curve), _p(rw), C.c_void_p(out[k].ctypes.data))
    return out


def score(ref, spec, row, pose6, cfg):
    spec = np.ascontiguousarray(spec, dtype=np.complex64)
    row = np.ascontiguousarray(row, dtype=ROW_DTYPE).reshape(1)
    pose = _f32(pose6)
    o4 = np.zeros(4, dtype=np.float32)
    s = lib().orc_score(ref._h, _p(spec), _p(row), _p(pose), C.byref(cfg), _p(o4))
    return float(s), o4


def score_grad(ref, spec, row, pose6, cfg, ring_cut=0):
    """(score x100, out4, dnum[5], dB[3], jtj[15]) — analytic derivatives of SEMANTICS.md §7c; ring_cut > 0 keeps the
    rings <= ring_cut only (coarse-to-fine stages)."""
    spec = np.ascontiguousarray(spec, dtype=np.complex64)
    row = np.ascontiguousarray(row, dtype=ROW_DTYPE).reshape(1)
    pose = _f32(pose6)
    o4, dn, db, jj = np.zeros(4, np.float32), np.zeros(5, np.float32), np.zeros(3, np.float32), np.zeros(15, np.float32)
    s = lib().orc_score_grad_cut(ref._h, _p(spec), _p(row), _p(pose), C.byref(cfg), _p(o4), _p(dn), _p(db), _p(jj), int(ring_cut))
    return float(s), o4, dn, db, jj


def normalize(img, radius_px, normalize=1, invert=0):
    img = _f32(img)
    out = np.zeros_like(img)
    lib().orc_normalize(_p(img), img.shape[-1], float(radius_px), int(normalize), int(invert), _p(out))
    return out


def phase_sum(ref, specs, rows, cfg):
    specs = np.ascontiguousarray(specs, dtype=np.complex64)
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    out = np.zeros((cfg.box, cfg.box // 2 + 1), dtype=np.complex64)
    lib().orc_phase_sum(ref._h, _p(specs), _p(rows), rows.size, C.byref(cfg), _p(out))
    return out


def focus_center(cfg, pose6):
    pose = _f32(pose6)
    cx, cy = np.zeros(1, np.float32), np.zeros(1, np.float32)
    lib().orc_focus_center(C.byref(cfg), _p(pose), _p(cx), _p(cy))
    return float(cx[0]), float(cy[0])


def focus_logp(ref, spec, row, pose6, cfg, o4):
    spec = np.ascontiguousarray(spec, dtype=np.complex64)
    row = np.ascontiguousarray(row, dtype=ROW_DTYPE).reshape(1)
    pose, o4 = _f32(pose6), _f32(o4)
    return float(lib().orc_focus_logp(ref._h, _p(spec), _p(row), _p(pose), C.byref(cfg), _p(o4)))


def refine_local(ref, specs, rows, cfg):
    specs = np.ascontiguousarray(specs, dtype=np.complex64)
    rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
    ne = lib().orc_refine_local(ref._h, _p(specs), _p(rows), rows.size, C.byref(cfg))
    return rows, int(ne)


def refine_local_fast(ref, specs, rows, cfg):
    """refine_local through the optimised CPU leg (cspb_oracle_fast.c): same algorithm, same results."""
    specs = np.ascontiguousarray(specs, dtype=np.complex64)
    rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
    ne = lib().orc_refine_local_fast(ref._h, _p(specs), _p(rows), rows.size, C.byref(cfg))
    return rows, int(ne)


def global_search(ref, specs, rows, cfg, angles3):
    specs = np.ascontiguousarray(specs, dtype=np.complex64)
    rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
    ang = _f32(angles3).reshape(-1, 3)
    ne = lib().orc_global_search(ref._h, _p(specs), _p(rows), rows.size, C.byref(cfg), _p(ang), ang.shape[0])
    return rows, int(ne)


def csp_compose(particle, particle0, tilt, tilt0, centre3, pixel, base_xy):
    """Projection pose (psi, theta, phi, x, y) from extended-table entries (numpy records)."""
    p = np.ascontiguousarray(particle, dtype=PARTICLE_DTYPE).reshape(1)
    p0 = np.ascontiguousarray(particle0, dtype=PARTICLE_DTYPE).reshape(1)
    t = np.ascontiguousarray(tilt, dtype=TILT_DTYPE).reshape(1)
    t0 = np.ascontiguousarray(tilt0, dtype=TILT_DTYPE).reshape(1)
    c3 = _f32(centre3)
    out = np.zeros(5, dtype=np.float32)
    lib().orc_csp_compose(_p(p), _p(p0), _p(t), _p(t0), _p(c3), float(pixel), float(base_xy[0]), float(base_xy[1]), _p(out))
    return out


def csp_run(ref, specs, rows, particles, tilts, cfg, csp_cfg, first, last):
    """external/CSP/csp numerics on the CPU; returns (rows, particles, tilts, n_evals)."""
    specs = np.ascontiguousarray(specs, dtype=np.complex64)
    rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
    particles = np.array(particles, dtype=PARTICLE_DTYPE, copy=True)
    tilts = np.array(tilts, dtype=TILT_DTYPE, copy=True)
    ne = lib().orc_csp_run(ref._h, _p(specs), _p(rows), rows.size, _p(particles), particles.size, _p(tilts), tilts.size,
                           C.byref(cfg), C.byref(csp_cfg), int(first), int(last))
    if ne < 0:
        raise ValueError(f"orc_csp_run failed ({ne})")
    return rows, particles, tilts, int(ne)


class Recon:
    def __init__(self, cfg):
        self.cfg = cfg
        self._h = lib().orc_recon_create(C.byref(cfg))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_recon_finish_fast(self._h)  # releases a pending raw accumulator pair of insert_fast
            lib().orc_recon_free(self._h)
            self._h = None

    def insert(self, imgs, rows, sym=None, aux=None):
        """aux: optional (n, 2) float32 {weight, cut radius in Fourier pixels} per row (dose weighting, SEMANTICS.md §10)."""
        imgs = _f32(imgs)
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        sym = None if sym is None else _f32(sym)
        aux = None if aux is None else _f32(aux).reshape(rows.size, 2)
        lib().orc_recon_insert_weighted(self._h, _p(imgs), _p(rows), rows.size, _p(sym), 0 if sym is None else sym.shape[0], _p(aux))

    def insert_fast(self, imgs, rows, sym=None, finish=True):
        """insert through the optimised CPU leg: only the coset representatives of the lattice-preserving symmetry
        subgroup are inserted per sample; `finish` (or finish_fast later) applies the lattice operators once to the sums."""
        imgs = _f32(imgs)
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        sym = None if sym is None else _f32(sym)
        lib().orc_recon_insert_fast(self._h, _p(imgs), _p(rows), rows.size, _p(sym), 0 if sym is None else sym.shape[0], 1 if finish else 0)

    def finish_fast(self):
        lib().orc_recon_finish_fast(self._h)

    def discard_fast(self):
        lib().orc_recon_discard_fast(self._h)

    def dump(self, half):
        npad = self.cfg.box * self.cfg.pad
        out = np.zeros((npad, npad, npad // 2 + 1, 4), dtype=np.float32)
        lib().orc_recon_get_dump(self._h, half, _p(out))
        return out

    def finalize(self, mw=0.0, outer_radius=0.0):
        n = self.cfg.box
        vol = np.zeros((n, n, n), dtype=np.float32)
        h1 = np.zeros_like(vol)
        h2 = np.zeros_like(vol)
        stats = np.zeros((n // 2 + 1, 7), dtype=np.float32)
        lib().orc_recon_finalize(self._h, mw, outer_radius, _p(h1), _p(h2), _p(vol), _p(stats))
        return vol, h1, h2, stats


def fsc(a, b):
    a, b = _f32(a), _f32(b)
    n = a.shape[0]
    out = np.zeros(n // 2 + 1, dtype=np.float32)
    lib().orc_fsc(_p(a), _p(b), n, _p(out))
    return out
