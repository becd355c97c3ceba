"""The band plan of the scorer (pyp_b200/csrc/plan.cu) is host code: compile it with a small driver and check
that every lattice sample of the band (SURVEY.md §8d counts) sits in exactly one slot, on the track of its
ring, for both ring orders."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_band_plan_covers_the_band_once(tmp_path):
    exe = str(tmp_path / "plan_check")
    src = os.path.join(ROOT, "pyp_b200", "csrc")
    subprocess.check_call([NVCC, "-std=c++17", "-O1", "-I", src, "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "plan_check.cu"), os.path.join(src, "plan.cu")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    env = {k: v for k, v in os.environ.items() if k != "CSPB_BAND_ORDER"}
    rows = [tuple(int(v) for v in line.split()) for line in subprocess.check_output([exe], env=env, text=True).splitlines()]
    assert len(rows) == 8
    want_band = {128: 4168, 256: 16558, 384: 37174, 512: 66031}      # SURVEY.md §8d for 128 / 256 / 384
    slots = {}
    for n, n_band, real, n_slots, n_bands, bad, radial in rows:
        assert bad == 0 and n_band == real == want_band[n]
        assert n_slots % 32 == 0 and n_band <= n_slots
        slots[(n, radial)] = n_slots
    for n in want_band:
        assert slots[(n, 0)] < slots[(n, 1)]                          # count-sorted bands pad less than radial ones
    assert slots[(256, 0)] == 17440 and slots[(256, 1)] == 18176      # 5.3 % vs 9.8 % padding
    assert slots[(128, 0)] <= 1.13 * want_band[128]
