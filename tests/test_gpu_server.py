"""The resident engine (pyp_b200/server.py) behind the drop-in executables: with CSPB_SERVER=auto the first bin/refine3d
call starts one daemon for the GPU, later calls (and concurrent ones) are served by it — same files as the stand-alone
processes, without a CUDA context, a reference transform or a stack read per call."""
import json
import os
import socket
import subprocess
import time

import numpy as np
import pytest

from common import small_case
from oracle import tier_a
from pyp_b200 import server, synth
from pyp_b200.formats import cistem, dump, mrc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def _call(prog, answers, cwd, env):
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(BIN, prog)], input=tier_a.heredoc(answers).encode(), cwd=cwd, env=env, capture_output=True, timeout=600)
    return r, time.perf_counter() - t0


def test_resident_engine_serves_the_drop_in_chain(tmp_path):
    n, px, P = 64, 1.35, 240
    ph, vol, rows, stack = small_case(n=n, n_part=P, snr=0.5)
    start = synth.perturb_rows(rows, 2.0, 1.0)
    dirs = {}
    for tag in ("alone", "served"):
        d = str(tmp_path / tag)
        os.makedirs(d)
        mrc.write(f"{d}/ds_stack.mrc", stack, px)
        mrc.write(f"{d}/ds_r01.mrc", vol, px)
        cistem.write_parameters(f"{d}/ds_r01.cistem", start)
        open(f"{d}/statistics_r01.txt", "w").close()
        dirs[tag] = d
    sockdir = str(tmp_path / "sock")
    os.makedirs(sockdir)
    base = dict(os.environ, CSPB_DEVICE="0", CSPB_SOCKET_DIR=sockdir)
    base.pop("CSPB_SERVER", None)
    env_srv = dict(base, CSPB_SERVER="auto", CSPB_SERVER_IDLE="120")
    ranges = [(1 + 30 * k, 30 * (k + 1)) for k in range(8)]

    def refine_answers(f, l):
        return tier_a.refine3d_answers("ds_stack.mrc", "ds_r01.cistem", "ds_r01.mrc", "statistics_r01.txt", "ds_r01", f, l, px, 100.0, 0.38 * n * px, 60.0, 4 * px)

    try:
        # stand-alone processes (one context each)
        t_alone = []
        for f, l in ranges[:2]:
            r, dt = _call("refine3d", refine_answers(f, l), dirs["alone"], base)
            assert r.returncode == 0, r.stderr
            t_alone.append(dt)
        # first served call starts the daemon; the second one finds context and reference in place
        r, dt_first = _call("refine3d", refine_answers(*ranges[0]), dirs["served"], env_srv)
        assert r.returncode == 0 and b"Refine3D: Normal termination" in r.stdout, r.stderr
        r, dt_second = _call("refine3d", refine_answers(*ranges[1]), dirs["served"], env_srv)
        assert r.returncode == 0 and b"Reference transform reused" in r.stdout, r.stderr
        for f, l in ranges[:2]:
            name = "ds_r01_%07d_%07d.cistem" % (f, l)
            assert open(f"{dirs['alone']}/{name}", "rb").read() == open(f"{dirs['served']}/{name}", "rb").read()
        assert dt_second < min(t_alone), (dt_second, t_alone)  # no context, no reference transform
        # the remaining ranges concurrently, as pyp's joblib fan-out would (mpi.py:44-48)
        procs = [subprocess.Popen([os.path.join(BIN, "refine3d")], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                  cwd=dirs["served"], env=env_srv) for _ in ranges[2:]]
        for p, (f, l) in zip(procs, ranges[2:]):
            p.stdin.write(tier_a.heredoc(refine_answers(f, l)).encode())
            p.stdin.close()
        for p in procs:
            assert p.wait(timeout=600) == 0, p.stderr.read()
        served = cistem.merge([f"{dirs['served']}/ds_r01_%07d_%07d.cistem" % r_ for r_ in ranges])
        assert list(served["position_in_stack"]) == list(range(1, P + 1))
        # reconstruct3d through the daemon reuses the stack ranges refine3d uploaded; accumulators equal the stand-alone ones
        cistem.write_parameters(f"{dirs['served']}/ds_r01_used.cistem", served)
        cistem.write_parameters(f"{dirs['alone']}/ds_r01_used.cistem", served)
        ans = tier_a.reconstruct3d_answers("ds_stack.mrc", "ds_r01_used.cistem", "ds_r01.mrc", "ds_r01", 1, P, px, 100.0, px * n / 2, 2 * px,
                                           "ds_r01_map1_n1.mrc", "ds_r01_map2_n1.mrc")
        r, _ = _call("reconstruct3d", ans, dirs["served"], env_srv)
        assert r.returncode == 0 and b"caught" not in r.stdout + r.stderr, r.stderr
        r, _ = _call("reconstruct3d", ans, dirs["alone"], base)
        assert r.returncode == 0, r.stderr
        for h in (1, 2):
            a, b = dump.read(f"{dirs['alone']}/ds_r01_map{h}_n1.mrc")[1], dump.read(f"{dirs['served']}/ds_r01_map{h}_n1.mrc")[1]
            assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()
        # errors come back like a stand-alone front-end's: non-zero status and the word pyp greps for
        bad = refine_answers(1, 30)
        bad[1] = "missing.cistem"
        r, _ = _call("refine3d", bad, dirs["served"], env_srv)
        assert r.returncode == 1 and b"caught" in r.stderr
        print(f"refine3d per range of 30 particles: stand-alone {min(t_alone):.2f} s, first served call {dt_first:.2f} s, served {dt_second:.2f} s")
    finally:
        try:
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.connect(os.path.join(sockdir, os.path.basename(server.socket_path(0))))
            s.sendall(b'{"prog": "shutdown"}\n')
            info = s.makefile("rb").readline()
            assert b"served" in info
        except OSError:
            pass
