"""The C-ABI library loads without a GPU and exports every symbol include/cspb200.h declares;
without a device every entry fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from pyp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cspb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cspb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cspb200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes prototypes out of sync with the header"


def test_struct_layouts_match_header():
    assert _lib.ROW_DTYPE.itemsize == 128
    assert ctypes.sizeof(_lib.RefineCfg) == 4 * (29 + 7)
    assert ctypes.sizeof(_lib.ReconCfg) == 4 * (12 + 8)
    lib = _lib.load()
    assert lib.cspb_abi_version() == 1
    cfg = _lib.RefineCfg()
    assert lib.cspb_refine_cfg_default(ctypes.byref(cfg), 128, 1.35) == 0
    assert cfg.box == 128 and cfg.pad == 1 and cfg.low_res_limit == 100.0 and cfg.signed_cc_limit == 30.0
    assert cfg.local_iterations == 8 and cfg.refine_psi == 1
    rc = _lib.ReconCfg()
    assert lib.cspb_recon_cfg_default(ctypes.byref(rc), 128, 1.35) == 0
    assert rc.resolution_limit == pytest.approx(2.7) and rc.mask_radius == pytest.approx(1.35 * 64)
    assert lib.cspb_refine_cfg_default(None, 128, 1.35) != 0
    # csp: extended-table rows and the csp_* config (defaults of config/pyp_config.toml [tabs.csp])
    assert _lib.PARTICLE_DTYPE.itemsize == 48 and _lib.TILT_DTYPE.itemsize == 24
    assert ctypes.sizeof(_lib.CspCfg) == 4 * (18 + 6)
    cc = _lib.CspCfg()
    assert lib.cspb_csp_cfg_default(ctypes.byref(cc)) == 0
    assert (cc.window_min, cc.window_max, cc.iterations, cc.random_evals) == (0, 20, 5, 0)
    assert cc.tol_particle_psi == 30.0 and cc.tol_tilt_angle == 1.5 and cc.tol_tilt_shift == 100.0 and cc.tol_defocus == 750.0
    # the host-side pose composition needs no GPU
    p = np.zeros(1, _lib.PARTICLE_DTYPE)
    p["psi"], p["theta"], p["phi"] = -30.0, -40.0, -50.0
    t = np.zeros(1, _lib.TILT_DTYPE)
    out = np.zeros(5, np.float32)
    c3 = np.zeros(3, np.float32)
    assert lib.cspb_csp_compose(_lib.ptr(p), _lib.ptr(p), _lib.ptr(t), _lib.ptr(t), _lib.ptr(c3), 1.0, 0.0, 0.0, _lib.ptr(out)) == 0
    assert np.allclose(out[:3], [30.0, 40.0, 50.0], atol=1e-3)   # SPA limit: particle = minus the projection angles


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.cspb_create(0, ctypes.byref(h)) == -2  # CSPB_E_CUDA
    from pyp_b200.engine import CspbError, Engine

    with pytest.raises(CspbError):
        Engine(0)


def test_product_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "pyp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|cspb_oracle\.h", text, flags=re.M):
                    bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"
