#!/usr/bin/env python
"""Batched 2-D R2C FFT: the hand-written sm_100a kernels (cspb_fft2_r2c) against cuFFT
(cspb_cufft2_r2c) on device-resident stacks, CUDA events on the engine stream.
Algorithmic bytes = 4 n^2 (read) + 8 n (n/2+1) (write) per image."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from pyp_b200.engine import Engine  # noqa: E402
from pyp_b200._lib import DEVICE  # noqa: E402


def main():
    eng = Engine(0)
    ext = torch.cuda.ExternalStream(eng.stream)
    out = []
    for n in (128, 256, 384, 512):
        batch = max(64, int(1.5e9 / (4 * n * n)))
        x = torch.randn(batch, n, n, device="cuda")
        y = torch.empty(batch, n, n // 2 + 1, 2, device="cuda")
        torch.cuda.synchronize()
        res = {}
        for name, fn in (("ours", eng._l.cspb_fft2_r2c), ("cufft", eng._l.cspb_cufft2_r2c)):
            for _ in range(2):
                eng._ck(fn(eng._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n, batch, DEVICE))
            eng.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            reps = 5
            for _ in range(reps):
                eng._ck(fn(eng._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n, batch, DEVICE))
            e1.record(ext)
            eng.sync()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gb = batch * (4 * n * n + 8 * n * (n // 2 + 1)) / 1e9
            res[name] = {"ms": ms, "us_per_image": 1e3 * ms / batch, "GBps": gb / (ms * 1e-3)}
        out.append({"n": n, "batch": batch, **res, "ours_over_cufft": res["cufft"]["ms"] / res["ours"]["ms"]})
        del x, y
    print(json.dumps(out))


if __name__ == "__main__":
    main()
