"""Python face of the C-ABI (include/cspb200.h): one :class:`Engine` = one GPU context.

This is the host-side mirror of what the reference reaches by spawning
``external/cistem2/{refine3d,reconstruct3d,local_merge3d,merge3d}``
(src/pyp/refine/frealign/frealign.py:3918-3994, 1780-1824, 1878-1888, 2075-2093): same
parameter tables (``.cistem`` rows as a numpy structured array, :data:`ROW_DTYPE`), same image
stacks (MRC mode-2 payload as float32 arrays), the numerics done by hand-written sm_100a kernels.
No CPU fallback exists; without the CUDA library / a B200 every call raises.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DEVICE, HOST, PARTICLE_DTYPE, ROW_DTYPE, TILT_DTYPE, CspbError, CspCfg, ReconCfg, RefineCfg, SelectCfg, ptr
from .symmetry import symmetry_matrices

__all__ = ["Engine", "ROW_DTYPE", "PARTICLE_DTYPE", "TILT_DTYPE", "RefineCfg", "ReconCfg", "CspCfg", "SelectCfg", "CspbError", "HOST", "DEVICE", "new_rows"]


def new_rows(n, pixel_size=1.0, voltage_kv=300.0, cs_mm=2.7, amplitude_contrast=0.07):
    """Fresh projection rows with the defaults pyp writes for new SPA particles
    (src/pyp/inout/metadata/core.py:1352-1374: angles/shifts 0, OCC 100, SIGMA 0.5, SCORE 0.5)."""
    rows = np.zeros(n, dtype=ROW_DTYPE)
    rows["position_in_stack"] = np.arange(1, n + 1, dtype=np.uint32)
    rows["occupancy"] = 100.0
    rows["sigma"] = 0.5
    rows["score"] = 0.5
    rows["pixel_size"] = pixel_size
    rows["voltage_kv"] = voltage_kv
    rows["cs_mm"] = cs_mm
    rows["amplitude_contrast"] = amplitude_contrast
    rows["image_is_active"] = 1
    return rows


def _loc_ptr(a):
    """(pointer, location) of a numpy array (host) or torch CUDA tensor (device).  The pointer holds no
    reference: the caller keeps `a` alive through the C call, so a numpy array must already be C-contiguous
    (a silent copy here would be freed before the library reads it)."""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "pass a C-contiguous array (np.ascontiguousarray bound to a local)"
        return ptr(a), HOST
    if hasattr(a, "data_ptr"):  # torch tensor
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr()), (DEVICE if a.is_cuda else HOST)
    raise TypeError(type(a))


class Engine:
    def __init__(self, device=0):
        self._l = _lib.lib()
        h = C.c_void_p()
        rc = self._l.cspb_create(int(device), C.byref(h))
        if rc != 0:
            raise CspbError(
                f"cspb_create(device={device}) failed with {rc}: no usable sm_100 GPU "
                "(libcspb200 has no CPU fallback)"
            )
        self._h = h
        self.device = int(device)
        self.box = None
        self._keep = []  # host arrays that must outlive async calls

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._l.cspb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self._l.cspb_last_error(self._h)
            raise CspbError(f"libcspb200 error {rc}: {msg.decode() if msg else ''}")

    def sync(self):
        self._ck(self._l.cspb_sync(self._h))

    @property
    def stream(self):
        s = C.c_void_p()
        self._ck(self._l.cspb_stream(self._h, C.byref(s)))
        return s.value or 0

    @property
    def launches(self):
        return int(self._l.cspb_launch_count(self._h))

    def profile_enable(self, on=True):
        self._ck(self._l.cspb_profile_enable(self._h, 1 if on else 0))

    def profile_get(self, kind):
        """(total_ms, launches, units) of kind 0 = scoring kernel, 1 = insertion kernel."""
        t, n, u = C.c_double(), C.c_int64(), C.c_int64()
        self._ck(self._l.cspb_profile_get(self._h, int(kind), C.byref(t), C.byref(n), C.byref(u)))
        return t.value, int(n.value), int(u.value)

    def count_loads(self, on=True):
        """Switch the scorer's gather-load census on / off (roofline bookkeeping, untimed steps only)."""
        self._ck(self._l.cspb_profile_count_loads(self._h, 1 if on else 0))

    def loads(self):
        """(32-byte reference loads issued by the scorer, 8-byte image slot reads, evaluations covered) since the
        census was switched on."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self._l.cspb_profile_get_loads(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    # ------------------------------------------------------------------ refine3d
    @staticmethod
    def refine_defaults(box, pixel_size):
        cfg = RefineCfg()
        rc = _lib.lib().cspb_refine_cfg_default(C.byref(cfg), int(box), float(pixel_size))
        if rc != 0:
            raise CspbError(f"cspb_refine_cfg_default failed: {rc}")
        return cfg

    def refine_configure(self, cfg: RefineCfg):
        self._ref_key = None
        self._ck(self._l.cspb_refine_configure(self._h, C.byref(cfg)))
        self.box = cfg.box
        self.rcfg = cfg

    def refine_reset_images(self):
        """Forget images, whitening curve and ring weights; keep the configuration and the reference."""
        self._ck(self._l.cspb_refine_reset_images(self._h))

    def ensure_reference(self, cfg: RefineCfg, ref_path, loader):
        """Configure + set the reference unless the context already holds exactly this configuration and this
        reference file (same path, size and mtime): a resident engine then skips the 3-D FFT of the map and only
        forgets the previous call's images.  `loader()` returns the (n, n, n) float32 volume.  Returns True when
        the cached reference was reused."""
        import os

        st = os.stat(ref_path)
        key = (bytes(cfg), os.path.abspath(ref_path), st.st_size, st.st_mtime_ns)
        if getattr(self, "_ref_key", None) == key:
            self.refine_reset_images()
            self.rcfg = cfg
            return True
        self._ref_key = None
        self.refine_configure(cfg)
        vol = np.ascontiguousarray(loader(), dtype=np.float32)
        if vol.shape != (cfg.box,) * 3:
            raise ValueError(f"reference {ref_path} {vol.shape} does not match the {cfg.box}-pixel stack")
        self.set_reference(vol)
        self._ref_key = key
        return False

    def band_counts(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self._l.cspb_band_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_reference(self, vol):
        if isinstance(vol, np.ndarray):
            vol = np.ascontiguousarray(vol, dtype=np.float32)  # bound to a local: alive through the C call
        n = vol.shape[0]
        assert tuple(vol.shape) == (n, n, n)
        p, loc = _loc_ptr(vol)
        self._ck(self._l.cspb_set_reference(self._h, p, n, loc))

    def set_symmetry(self, symbol_or_mats):
        mats = symmetry_matrices(symbol_or_mats) if isinstance(symbol_or_mats, str) else np.ascontiguousarray(symbol_or_mats, dtype=np.float32)
        self._ck(self._l.cspb_set_symmetry(self._h, ptr(mats), int(mats.shape[0])))
        return int(mats.shape[0])

    def set_ring_weights(self, w):
        if w is None:
            self._ck(self._l.cspb_refine_set_ring_weights(self._h, None, 0))
        else:
            w = np.ascontiguousarray(w, dtype=np.float32)
            self._ck(self._l.cspb_refine_set_ring_weights(self._h, ptr(w), int(w.size)))

    def set_focus_mask(self, x, y, z, radius):
        """2-D focus mask of refine3d (answers 29-32 + 44; class_focusmask, frealign.py:3845-3848,3883-3885):
        sphere centre / radius in Angstrom from the corner of the map; radius <= 0 switches it off."""
        self._ck(self._l.cspb_refine_set_focus_mask(self._h, float(x), float(y), float(z), float(radius)))

    def phase_sum(self, rows):
        """Beam-tilt input of refine_ctf: sum of G * conj(CTF * slice) over the loaded images at the poses of
        `rows`, on the (box, box/2+1) half-plane grid (zero outside the scoring band)."""
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        out = np.zeros((self.box, self.box // 2 + 1), dtype=np.complex64)
        self._ck(self._l.cspb_refine_phase_sum(self._h, ptr(rows), int(rows.size), ptr(out)))
        return out

    def noise_curve(self):
        out = np.zeros(self.box + 1, dtype=np.float32)
        self._ck(self._l.cspb_refine_get_noise_curve(self._h, ptr(out), out.size))
        return out

    def set_noise_curve(self, curve):
        curve = np.ascontiguousarray(curve, dtype=np.float32)
        self._ck(self._l.cspb_refine_set_noise_curve(self._h, ptr(curve), curve.size))

    def load_images(self, images, append=False):
        if isinstance(images, np.ndarray):
            images = np.ascontiguousarray(images, dtype=np.float32)
        p, loc = _loc_ptr(images)
        n_img = int(images.shape[0])
        assert tuple(images.shape[1:]) == (self.box, self.box)
        self._ck(self._l.cspb_refine_load_images(self._h, p, n_img, loc, 1 if append else 0))

    def keep_spectra(self, on=True):
        """One forward transform per projection for refinement and insertion (cspb_refine_keep_spectra): the device tensor
        passed to load_images must reach recon_insert unchanged."""
        self._ck(self._l.cspb_refine_keep_spectra(self._h, 1 if on else 0))

    @property
    def num_images(self):
        return int(self._l.cspb_refine_num_images(self._h))

    def score(self, rows):
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        out = np.zeros(rows.size, dtype=np.float32)
        self._ck(self._l.cspb_refine_score(self._h, ptr(rows), rows.size, ptr(out)))
        return out

    def score_poses(self, rows, image_index, poses6):
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        idx = np.ascontiguousarray(image_index, dtype=np.int32)
        poses = np.ascontiguousarray(poses6, dtype=np.float32).reshape(-1, 6)
        assert poses.shape[0] == idx.size
        out = np.zeros(idx.size, dtype=np.float32)
        self._ck(self._l.cspb_refine_score_poses(self._h, ptr(rows), rows.size, ptr(idx), ptr(poses), idx.size, ptr(out)))
        return out

    def score_grad(self, rows, ring_cut=0):
        """Value + analytic derivatives at the poses of `rows`: (n, 28) array {num, X, A, B, dnum[5], dB[3], jtj[15], 0}."""
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        out = np.zeros((rows.size, 28), dtype=np.float32)
        self._ck(self._l.cspb_refine_score_grad(self._h, ptr(rows), rows.size, int(ring_cut), ptr(out)))
        return out

    def matching_projections(self, rows):
        """CTF-multiplied projections of the reference at the poses of `rows`, in the frame of the particle images
        (refine3d's `_match.mrc_<range>` stack): (n_rows, box, box) float32."""
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        out = np.zeros((rows.size, self.box, self.box), dtype=np.float32)
        self._ck(self._l.cspb_refine_matching_projections(self._h, ptr(rows), rows.size, ptr(out)))
        return out

    def set_search_grid(self, angles3):
        """Global-search orientation grid: (n, 3) array of (psi, theta, phi) in degrees."""
        a = np.ascontiguousarray(angles3, dtype=np.float32).reshape(-1, 3)
        self._ck(self._l.cspb_refine_set_search_grid(self._h, ptr(a), int(a.shape[0])))
        return int(a.shape[0])

    def refine(self, rows, want_changes=False):
        """Run refine3d's search over the loaded images.  Returns (rows, changes|None, n_evals)."""
        rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
        changes = np.zeros_like(rows) if want_changes else None
        ne = C.c_int64(0)
        self._ck(self._l.cspb_refine_run(self._h, ptr(rows), rows.size, ptr(changes), C.byref(ne)))
        return rows, changes, int(ne.value)

    def refine_device(self, rows_dev_ptr, n):
        """Enqueue the refinement with rows resident on the device (benchmark path)."""
        ne = C.c_int64(0)
        self._ck(self._l.cspb_refine_run_device(self._h, C.c_void_p(int(rows_dev_ptr)), int(n), C.byref(ne)))
        return int(ne.value)

    # ------------------------------------------------------------------ streamed host pipeline
    DO_REFINE, DO_INSERT = 1, 2

    def refine_reconstruct(self, images, rows, refine=True, insert=True):
        """refine3d and/or reconstruct3d insertion over a host stack with one upload per projection,
        copies overlapped with compute (pin the stack for full overlap).  Returns (rows, n_evals)."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
        assert images.shape[0] == rows.size
        ne = C.c_int64(0)
        flags = (self.DO_REFINE if refine else 0) | (self.DO_INSERT if insert else 0)
        self._ck(self._l.cspb_refine_reconstruct(self._h, ptr(images), ptr(rows), rows.size, flags, C.byref(ne)))
        return rows, int(ne.value)

    # ------------------------------------------------------------------ between the stages: score shaping, occupancies
    @staticmethod
    def select_defaults(cutoff=1.0):
        cfg = SelectCfg()
        rc = _lib.lib().cspb_select_cfg_default(C.byref(cfg))
        if rc != 0:
            raise CspbError(f"cspb_select_cfg_default failed: {rc}")
        cfg.cutoff = float(cutoff)
        return cfg

    def select_scores(self, rows, cfg: SelectCfg, tilt_angle=None):
        """shape_phase_residuals on the device (scores.py:300-761): returns (shaped copy of rows, threshold)."""
        rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
        tilt = None if tilt_angle is None else np.ascontiguousarray(tilt_angle, dtype=np.float32)
        thr = C.c_double(float("nan"))
        self._ck(self._l.cspb_select_scores(self._h, ptr(rows), rows.size, ptr(tilt), C.byref(cfg), HOST, C.byref(thr)))
        return rows, thr.value

    def class_occupancies(self, logp, sigma, class_average_occ):
        """occupancy_extended on the device (occupancies.py:173-208): (occ (K, n) percent, sigma (n,))."""
        logp = np.ascontiguousarray(logp, dtype=np.float32)
        sigma = np.ascontiguousarray(sigma, dtype=np.float32)
        avg = np.ascontiguousarray(class_average_occ, dtype=np.float64)
        K, n = logp.shape
        occ, sg = np.zeros((K, n), dtype=np.float32), np.zeros(n, dtype=np.float32)
        self._ck(self._l.cspb_class_occupancies(self._h, ptr(logp), ptr(sigma), ptr(avg), K, n, ptr(occ), ptr(sg), HOST))
        return occ, sg

    def global_weights(self, rows, n_idx=None):
        """Mean SCORE per scan-order index over the projections with occupancy > 0, -1 where none (pyp's global_weight.txt,
        inout/metadata/core.py:3039-3075) — on the device; equals tables.global_weights."""
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        n_idx = int(n_idx or (int(rows["tind"].max()) + 1 if rows.size else 1))
        out, used = np.zeros(n_idx, dtype=np.float64), C.c_int(0)
        self._ck(self._l.cspb_global_weights(self._h, ptr(rows), rows.size, HOST, ptr(out), n_idx, C.byref(used)))
        return out[: used.value]

    def refine_select_reconstruct(self, images, rows, cfg: SelectCfg):
        """refine3d -> score shaping -> reconstruct3d insertion over a host stack without leaving the device.
        Returns (refined + shaped rows, n_evals, threshold)."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
        assert images.shape[0] == rows.size
        ne, thr = C.c_int64(0), C.c_double(float("nan"))
        self._ck(self._l.cspb_refine_select_reconstruct(self._h, ptr(images), ptr(rows), rows.size, C.byref(cfg), C.byref(ne), C.byref(thr)))
        return rows, int(ne.value), thr.value

    # ------------------------------------------------------------------ csp (external/CSP/csp)
    @staticmethod
    def csp_defaults(mode=5):
        cfg = CspCfg()
        rc = _lib.lib().cspb_csp_cfg_default(C.byref(cfg))
        if rc != 0:
            raise CspbError(f"cspb_csp_cfg_default failed: {rc}")
        cfg.mode = int(mode)
        return cfg

    def csp_run(self, rows, particles, tilts, cfg: CspCfg, first=0, last=-1):
        """Constrained refinement of particles / tilts first..last over the loaded images
        (rows[k] <-> image k).  Returns (rows, particles, tilts, n_evals), inputs untouched."""
        rows = np.array(rows, dtype=ROW_DTYPE, copy=True)
        particles = np.array(particles, dtype=PARTICLE_DTYPE, copy=True)
        tilts = np.array(tilts, dtype=TILT_DTYPE, copy=True)
        ne = C.c_int64(0)
        self._ck(self._l.cspb_csp_run(self._h, ptr(rows), rows.size, ptr(particles), particles.size, ptr(tilts), tilts.size,
                                      C.byref(cfg), int(first), int(last), C.byref(ne)))
        return rows, particles, tilts, int(ne.value)

    @staticmethod
    def csp_compose(particle, particle0, tilt, tilt0, centre3, pixel_size, base_xy):
        p = np.ascontiguousarray(particle, dtype=PARTICLE_DTYPE).reshape(1)
        p0 = np.ascontiguousarray(particle0, dtype=PARTICLE_DTYPE).reshape(1)
        t = np.ascontiguousarray(tilt, dtype=TILT_DTYPE).reshape(1)
        t0 = np.ascontiguousarray(tilt0, dtype=TILT_DTYPE).reshape(1)
        c3 = np.ascontiguousarray(centre3, dtype=np.float32)
        out = np.zeros(5, dtype=np.float32)
        rc = _lib.lib().cspb_csp_compose(ptr(p), ptr(p0), ptr(t), ptr(t0), ptr(c3), float(pixel_size), float(base_xy[0]), float(base_xy[1]), ptr(out))
        if rc != 0:
            raise CspbError(f"cspb_csp_compose failed: {rc}")
        return out

    def csp_extract(self, images, rows, box_in, binning=1):
        """csp mode -2: cut (and bin) the particle boxes of `rows` out of a tilt series
        (n_tilt, ny, nx) float32; returns the stack (n_rows, box_in/bin, box_in/bin).  With a CUDA tensor as `images`
        the stack is returned as a CUDA tensor: tilt series -> boxes -> load_images without a stack on the host (the
        54 GB stack of BASELINE configs[2] never has to exist)."""
        if hasattr(images, "is_cuda") and images.is_cuda:
            import torch

            rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
            nt, ny, nx = (int(v) for v in images.shape)
            bo = int(box_in) // int(binning)
            out = torch.empty((rows.size, bo, bo), dtype=torch.float32, device=images.device)
            torch.cuda.current_stream(images.device).synchronize()
            self._ck(self._l.cspb_csp_extract(self._h, C.c_void_p(images.data_ptr()), nx, ny, nt, ptr(rows), rows.size, int(box_in), int(binning),
                                              C.c_void_p(out.data_ptr()), DEVICE))
            return out
        images = np.ascontiguousarray(images, dtype=np.float32)
        rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
        nt, ny, nx = images.shape
        bo = int(box_in) // int(binning)
        out = np.zeros((rows.size, bo, bo), dtype=np.float32)
        self._ck(self._l.cspb_csp_extract(self._h, ptr(images), nx, ny, nt, ptr(rows), rows.size, int(box_in), int(binning), ptr(out), HOST))
        return out

    def spa_extract(self, micrograph, coords_xy, box, coordinate_binning=1.0, to_device=False):
        """Box cutting of extract_particles_non_mpi (extract/core.py:360-511) from one micrograph (ny, nx) float32 at the
        (x, y) coordinates `coords_xy` (n, 2).  Returns the (n, box, box) stack — a CUDA tensor with `to_device` (the boxes
        then go to load_images / recon_insert without ever existing on the host)."""
        mic = np.ascontiguousarray(micrograph, dtype=np.float32)
        xy = np.ascontiguousarray(coords_xy, dtype=np.float32).reshape(-1, 2)
        ny, nx = mic.shape
        n = xy.shape[0]
        if to_device:
            import torch

            out = torch.empty((n, int(box), int(box)), dtype=torch.float32, device=torch.device("cuda", self.device))
            torch.cuda.current_stream(out.device).synchronize()
            self._ck(self._l.cspb_spa_extract(self._h, ptr(mic), nx, ny, ptr(xy), n, int(box), float(coordinate_binning), C.c_void_p(out.data_ptr()), HOST, DEVICE))
            return out
        out = np.zeros((n, int(box), int(box)), dtype=np.float32)
        self._ck(self._l.cspb_spa_extract(self._h, ptr(mic), nx, ny, ptr(xy), n, int(box), float(coordinate_binning), ptr(out), HOST, HOST))
        return out

    # ------------------------------------------------------------------ reconstruct3d / merge3d
    @staticmethod
    def recon_defaults(box, pixel_size):
        cfg = ReconCfg()
        rc = _lib.lib().cspb_recon_cfg_default(C.byref(cfg), int(box), float(pixel_size))
        if rc != 0:
            raise CspbError(f"cspb_recon_cfg_default failed: {rc}")
        return cfg

    def recon_begin(self, cfg: ReconCfg):
        self._ck(self._l.cspb_recon_begin(self._h, C.byref(cfg)))
        self.ccfg = cfg

    def recon_insert(self, images, rows, weight_cut=None):
        """weight_cut: optional (n, 2) float32 {weight, cut radius in Fourier pixels} per projection — the data-driven
        dose weighting of reconstruct3d's prompt 22 (tables.dose_weight_pairs)."""
        if isinstance(images, np.ndarray):
            images = np.ascontiguousarray(images, dtype=np.float32)
            rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
            wc = None if weight_cut is None else np.ascontiguousarray(weight_cut, dtype=np.float32).reshape(rows.size, 2)
            self._ck(self._l.cspb_recon_insert_weighted(self._h, ptr(images), ptr(rows), int(images.shape[0]), HOST, ptr(wc)))
        else:  # torch CUDA tensors; rows = device pointer (int), uint8 tensor of packed rows, or a host table (uploaded here)
            keep = None
            if isinstance(rows, np.ndarray):
                import torch

                keep = torch.from_numpy(np.ascontiguousarray(rows, dtype=ROW_DTYPE).view(np.uint8).reshape(-1, 128)).to(images.device)
                torch.cuda.current_stream(images.device).synchronize()  # the engine runs on its own stream
                rows = keep
            rp = rows if isinstance(rows, int) else rows.data_ptr()
            keep_w = None
            if weight_cut is not None:
                import torch

                keep_w = torch.from_numpy(np.ascontiguousarray(weight_cut, dtype=np.float32).reshape(-1, 2)).to(images.device)
                torch.cuda.current_stream(images.device).synchronize()
            self._ck(self._l.cspb_recon_insert_weighted(self._h, C.c_void_p(images.data_ptr()), C.c_void_p(rp), int(images.shape[0]), DEVICE,
                                                        None if keep_w is None else C.c_void_p(keep_w.data_ptr())))
            if keep is not None or keep_w is not None:
                self.sync()

    def recon_dims(self):
        npad, nf = C.c_int(), C.c_int64()
        self._ck(self._l.cspb_recon_dims(self._h, C.byref(npad), C.byref(nf)))
        return npad.value, int(nf.value)

    def recon_device_ptr(self, half):
        p = C.c_void_p()
        self._ck(self._l.cspb_recon_device_ptr(self._h, int(half), C.byref(p)))
        return p.value

    def recon_get_dump(self, half):
        npad, nf = self.recon_dims()
        out = np.zeros((npad, npad, npad // 2 + 1, 4), dtype=np.float32)
        self._ck(self._l.cspb_recon_get_dump(self._h, int(half), ptr(out), HOST))
        return out

    def recon_add_dump(self, half, dump):
        dump = np.ascontiguousarray(dump, dtype=np.float32)
        npad, nf = self.recon_dims()
        assert dump.size == nf, "dump size does not match the accumulator"
        self._ck(self._l.cspb_recon_add_dump(self._h, int(half), ptr(dump), HOST))

    def recon_finalize(self, molecular_mass_kda=0.0, outer_radius=0.0, want_halves=True, out=None):
        """merge3d finalise into host arrays.  `out` = optional (map, half1, half2) float32 arrays of
        shape (n, n, n) to fill (e.g. pinned buffers that are reused between calls)."""
        n = self.ccfg.box
        ns = n // 2 + 1
        if out is not None:
            vol, h1, h2 = out
            for a in (vol, h1, h2):
                assert a is None or (a.dtype == np.float32 and a.shape == (n, n, n) and a.flags["C_CONTIGUOUS"])
        else:
            vol = np.zeros((n, n, n), dtype=np.float32)
            h1 = np.zeros_like(vol) if want_halves else None
            h2 = np.zeros_like(vol) if want_halves else None
        stats = np.zeros((ns, 7), dtype=np.float32)
        self._ck(self._l.cspb_recon_finalize(self._h, float(molecular_mass_kda), float(outer_radius), ptr(h1), ptr(h2), ptr(vol), ptr(stats), ns, HOST))
        return vol, h1, h2, stats

    def recon_finalize_device(self, out_maps, molecular_mass_kda=0.0, outer_radius=0.0):
        """merge3d finalise with the three output volumes (map, half1, half2) as CUDA tensors."""
        n = self.ccfg.box
        ns = n // 2 + 1
        stats = np.zeros((ns, 7), dtype=np.float32)
        m, h1, h2 = out_maps
        self._ck(self._l.cspb_recon_finalize(self._h, float(molecular_mass_kda), float(outer_radius), C.c_void_p(h1.data_ptr()), C.c_void_p(h2.data_ptr()), C.c_void_p(m.data_ptr()), ptr(stats), ns, DEVICE))
        return stats

    def recon_end(self):
        self._ck(self._l.cspb_recon_end(self._h))

    # ------------------------------------------------------------------ building blocks
    def fft2_r2c(self, images):
        images = np.ascontiguousarray(images, dtype=np.float32)
        b, n = images.shape[0], images.shape[1]
        out = np.zeros((b, n, n // 2 + 1), dtype=np.complex64)
        self._ck(self._l.cspb_fft2_r2c(self._h, ptr(images), ptr(out), n, b, HOST))
        return out

    def fft2_c2r(self, spec):
        spec = np.ascontiguousarray(spec, dtype=np.complex64)
        b, n = spec.shape[0], spec.shape[1]
        out = np.zeros((b, n, n), dtype=np.float32)
        self._ck(self._l.cspb_fft2_c2r(self._h, ptr(spec), ptr(out), n, b, HOST))
        return out

    def cufft2_r2c(self, images):
        images = np.ascontiguousarray(images, dtype=np.float32)
        b, n = images.shape[0], images.shape[1]
        out = np.zeros((b, n, n // 2 + 1), dtype=np.complex64)
        self._ck(self._l.cspb_cufft2_r2c(self._h, ptr(images), ptr(out), n, b, HOST))
        return out

    def ctf_image(self, row, n):
        row = np.ascontiguousarray(row, dtype=ROW_DTYPE).reshape(1)
        out = np.zeros((n, n // 2 + 1), dtype=np.float32)
        self._ck(self._l.cspb_ctf_image(self._h, ptr(row), int(n), ptr(out)))
        return out

    @property
    def wave_units(self):
        """Units of the scoring kernel resident at once on this GPU (one full wave)."""
        return int(self._l.cspb_wave_units(self._h))

    def gather_peak(self, window_bytes, per_cta=False):
        """GB/s of random 32-byte gathers from a window (roofline denominator of the scorer)."""
        g = C.c_float(0)
        self._ck(self._l.cspb_gather_peak(self._h, C.c_size_t(int(window_bytes)), 1 if per_cta else 0, C.byref(g)))
        return float(g.value)

    def project(self, psi, theta, phi):
        n = self.box
        out = np.zeros((n, n // 2 + 1), dtype=np.complex64)
        self._ck(self._l.cspb_project(self._h, float(psi), float(theta), float(phi), ptr(out)))
        return out
