#!/usr/bin/env python
"""Drive the drop-in executables at the benchmark's shape (256-px box, O symmetry) the way pyp does:
refine3d over R ranges, reconstruct3d over R ranges with dumps, merge3d.  Reports wall times (total and per range),
the pose error against the truth and the FSC = 0.143 crossing of the merged half maps.
Usage: cli_chain_check.py [particles=4096] [ranges=2]; with CSPB_SERVER=auto in the environment the executables are clients
of the resident per-GPU engine (pyp_b200/server.py), which is shut down at the end."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import angular_distance  # noqa: E402
from pyp_b200 import synth, synth_torch  # noqa: E402
from pyp_b200.formats import cistem, mrc, statistics  # noqa: E402

BIN = os.path.join(ROOT, "bin")


def sh(prog, answers, cwd, log):
    cmd = f"{BIN}/{prog} << eot >> {log} 2>&1\n" + "\n".join(str(a) for a in answers) + "\neot\n"
    t0 = time.perf_counter()
    rc = subprocess.run(cmd, shell=True, cwd=cwd, timeout=1200).returncode
    return rc, time.perf_counter() - t0


def main():
    n, px, P, sym = 256, 1.0, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, "O"
    dev = torch.device("cuda", 0)
    centres, amps, sigma = synth_torch.symmetric_phantom(n, sym)
    vol = synth_torch.volume(n, centres, amps, sigma, dev).cpu().numpy()
    rows = synth.make_rows(P, px, seed=1)
    stack = synth_torch.make_stack(n, centres, amps, sigma, rows, snr=0.05, seed=2, device=dev).cpu().numpy()
    start = synth.perturb_rows(rows, 2.0, 1.0)
    d = tempfile.mkdtemp()
    mrc.write(f"{d}/ds_stack.mrc", stack, px)
    mrc.write(f"{d}/ds_r01.mrc", vol, px)
    cistem.write_parameters(f"{d}/ds_r01.cistem", start)
    open(f"{d}/statistics_r01.txt", "w").close()
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    out = {"particles": P, "box": n, "symmetry": sym, "ranges": R, "resident_engine": os.environ.get("CSPB_SERVER", "")}
    outs, t_ref, t_each = [], 0.0, []
    edges = [round(k * P / R) for k in range(R + 1)]
    ranges = [(edges[k] + 1, edges[k + 1]) for k in range(R)]
    for first, last in ranges:
        ranger = "%07d_%07d" % (first, last)
        a = ["ds_stack.mrc", "ds_r01.cistem", "null", "ds_r01.mrc", "statistics_r01.txt", "no", "no", f"ds_r01_match.mrc_{ranger}",
             f"ds_r01_{ranger}.cistem", f"ds_r01_{ranger}_changes.cistem", sym, first, last, 1, px, 440.0, 0, 0.38 * n * px, 100.0, 2.5 * px,
             "30.0", 8.0, 1.5 * 0.38 * n * px, 8 * px, 20.0, 20, 0, 0, 0, 0, 0, 0, 500, "50.0", 1, "no", "yes", "yes", "yes", "yes", "yes",
             "yes", "no", "no", "no", "yes", "no", "no", "no", "no"]
        rc, dt = sh("refine3d", a, d, "refine.log")
        assert rc == 0, open(f"{d}/refine.log").read()[-2000:]
        t_ref += dt
        t_each.append(round(dt, 3))
        outs.append(f"{d}/ds_r01_{ranger}.cistem")
    refined = cistem.merge(outs)
    out["refine3d_seconds_all_ranges"] = t_ref
    out["refine3d_seconds_each_range"] = t_each
    out["median_angular_error_deg_start"] = float(np.median(angular_distance(start, rows)))
    out["median_angular_error_deg_refined"] = float(np.median(angular_distance(refined, rows)))
    cistem.write_parameters(f"{d}/ds_r01_used.cistem", refined)
    os.makedirs(f"{d}/scratch", exist_ok=True)
    t_rec, t_each = 0.0, []
    for k, (first, last) in enumerate(ranges, start=1):
        a = ["ds_stack.mrc", "ds_r01_used.cistem", "null", "ds_r01.mrc", "ds_r01_map1.mrc", "ds_r01_map2.mrc", "output.mrc", f"ds_r01_n{first}.res",
             sym, first, last, px, 440.0, 0, px * n / 2, 2 * px, 0, 2.0, "no", 0, -1, "no", 0, 1, 1, "yes", "no", "no", "no", "no", "yes", "no",
             "no", "no", "no", "yes", f"scratch/ds_r01_map1_n{k}.mrc", f"scratch/ds_r01_map2_n{k}.mrc", 1]
        rc, dt = sh("reconstruct3d", a, d, "recon.log")
        assert rc == 0 and "caught" not in open(f"{d}/recon.log").read()
        t_rec += dt
        t_each.append(round(dt, 3))
    out["reconstruct3d_seconds_all_ranges"] = t_rec
    out["reconstruct3d_seconds_each_range"] = t_each
    a = ["ds_r01_02_half1.mrc", "ds_r01_02_half2.mrc", "ds_r01_02.mrc", "ds_r01_02_statistics.txt", 440.0, 0, px * n / 2,
         "scratch/ds_r01_map1_n.mrc", "scratch/ds_r01_map2_n.mrc", R]
    rc, dt = sh("merge3d", a, d, "merge.log")
    log = open(f"{d}/merge.log").read()
    assert rc == 0 and "Merge3D: Normal termination" in log, log[-2000:]
    out["merge3d_seconds"] = dt
    out["merge3d_timing"] = [ln for ln in log.splitlines() if ln.startswith("timing:")]
    st = statistics.read_statistics(f"{d}/ds_r01_02_statistics.txt")
    fsc = st[:, 3]
    below = np.nonzero(fsc[1:] < 0.143)[0]
    out["fsc_0.143_resolution_A"] = float(st[1 + below[0], 1]) if below.size else float(st[-1, 1])
    _, rec = mrc.read(f"{d}/ds_r01_02.mrc")
    out["map_correlation_with_phantom"] = float(np.corrcoef(np.asarray(rec).ravel(), vol.ravel())[0, 1])
    print(json.dumps(out, indent=1))
    subprocess.run(["rm", "-rf", d])


def stop_daemons():
    """stop the resident engines the front-ends started (CSPB_SERVER=auto)"""
    import socket

    from pyp_b200 import server

    for g in range(torch.cuda.device_count()):
        try:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(server.socket_path(g))
            c.sendall(b'{"prog": "shutdown"}\n')
            c.makefile("rb").readline()
        except OSError:
            pass


if __name__ == "__main__":
    try:
        main()
    finally:
        if os.environ.get("CSPB_SERVER"):
            stop_daemons()
