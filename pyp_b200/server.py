"""Resident engine: one daemon per GPU behind the drop-in executables.

pyp launches `slurm_tasks` short-lived processes per stage, one per contiguous particle range
(src/pyp/system/local_run.py:507-516, frealign.py:3183).  A CUDA context per process costs ~1.4 s, the 3-D
transform of the reference map is repeated by every range, and refine3d / reconstruct3d each read the same stack
range from disk.  With `CSPB_SERVER=auto` (or `1`) the executables in bin/ become thin stdlib-only clients
(pyp_b200/cli/front.py) of this daemon, which keeps ONE context per GPU alive between invocations together with

  * the transformed reference (`Engine.ensure_reference`: reused while path, size, mtime and configuration match),
  * the particle stack in HBM (`StackCache`: ranges uploaded by refine3d are found again by reconstruct3d and by
    the next iteration; a 100 000 x 256 x 256 stack is 26 GB of the 180 GB),

and serves the requests of that GPU one at a time (the listen backlog is the queue for the concurrent callers).
Every request runs the SAME front-end code as a stand-alone invocation (`run(parse(answers), session=...)`), so the
files written and the log text returned are identical.

    python -m pyp_b200.server --device 0            # normally started by the first client (CSPB_SERVER=auto)

Protocol (Unix stream socket `$CSPB_SOCKET_DIR/cspb200-<uid>-gpu<k>.sock`): one request per connection, a JSON line
{"prog", "argv", "cwd", "stdin", "env"} answered by a JSON line {"rc", "out", "err"}; {"prog": "shutdown"} stops the daemon.
"""
import argparse
import contextlib
import io
import json
import os
import socket
import sys
import tempfile
import time
import traceback

import numpy as np


def socket_path(device):
    d = os.environ.get("CSPB_SOCKET_DIR") or tempfile.gettempdir()
    return os.path.join(d, f"cspb200-{os.getuid()}-gpu{int(device)}.sock")


class StackCache:
    """Device-resident particle stacks keyed by (path, size, mtime).  `__call__(path, positions)` returns the images at
    the 1-based stack positions as a CUDA tensor, uploading only the images not seen before."""

    def __init__(self, device, budget_bytes):
        import torch

        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.budget = budget_bytes
        self.stacks = {}   # key -> dict(data=tensor (nz, n, n), have=np.bool_[nz], mm=np.memmap, used=timestamp)
        self.uploaded_bytes = 0
        self.hit_bytes = 0

    def _entry(self, path):
        from .formats import mrc

        st = os.stat(path)
        key = (os.path.abspath(path), st.st_size, st.st_mtime_ns)
        e = self.stacks.get(key)
        if e is None:
            for k in [k for k in self.stacks if k[0] == key[0]]:  # the file changed: drop the stale copy
                del self.stacks[k]
            hdr, mm = mrc.read(path)  # float32 stacks (MRC mode 2, what pyp writes) are memory-mapped
            nz, ny, nx = hdr["nz"], hdr["ny"], hdr["nx"]
            need = nz * ny * nx * 4
            while self.stacks and sum(v["data"].numel() * 4 for v in self.stacks.values()) + need > self.budget:
                oldest = min(self.stacks, key=lambda k: self.stacks[k]["used"])
                del self.stacks[oldest]
            if need > self.budget:
                return None
            e = {"data": self.torch.empty((nz, ny, nx), dtype=self.torch.float32, device=self.dev), "have": np.zeros(nz, dtype=bool), "mm": mm}
            self.stacks[key] = e
        e["used"] = time.time()
        return e

    def __call__(self, path, positions):
        from .formats import mrc

        torch = self.torch
        pos = np.asarray(positions, dtype=np.int64) - 1
        e = self._entry(path)
        if e is None:  # does not fit the budget: plain read
            _, data = mrc.read(path, first=int(pos.min()) + 1, last=int(pos.max()) + 1)
            return np.ascontiguousarray(data[pos - pos.min()])
        missing = pos[~e["have"][pos]]
        if missing.size:
            missing = np.unique(missing)
            # contiguous runs of missing images: one host read + one upload each
            cuts = np.nonzero(np.diff(missing) != 1)[0] + 1
            for run in np.split(missing, cuts):
                a, b = int(run[0]), int(run[-1]) + 1
                host = torch.from_numpy(np.ascontiguousarray(e["mm"][a:b]))
                e["data"][a:b].copy_(host, non_blocking=False)
                self.uploaded_bytes += host.numel() * 4
            e["have"][missing] = True
        self.hit_bytes += (pos.size - missing.size) * e["data"][0].numel() * 4
        torch.cuda.current_stream(self.dev).synchronize()  # the engine runs on its own stream
        if pos.size and np.all(np.diff(pos) == 1):
            return e["data"][int(pos[0]):int(pos[-1]) + 1]
        return e["data"].index_select(0, torch.from_numpy(pos).to(self.dev)).contiguous()


def _dispatch(prog, argv, stdin_text, session, out):
    from .cli import csp, local_merge3d, merge3d, prompts, reconstruct3d, refine3d, refine_ctf

    mods = {"refine3d": refine3d, "reconstruct3d": reconstruct3d, "merge3d": merge3d, "local_merge3d": local_merge3d, "refine_ctf": refine_ctf}
    if prog in ("csp", "csp_GS"):
        return csp.main(argv, out=out, session=session)
    if prog not in mods:
        raise ValueError(f"unknown program {prog!r}")
    m = mods[prog]
    m.run(m.parse(prompts.Answers(stdin_text, prog)), out=out, session=session)
    return 0


def serve(device, idle_timeout=900.0, stack_cache_gb=64.0):
    from .cli.session import Session
    from .engine import Engine

    path = socket_path(device)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        os.unlink(path)
    except FileNotFoundError:
        pass
    eng = Engine(device)
    cache = StackCache(device, int(stack_cache_gb * (1 << 30))) if stack_cache_gb > 0 else None
    session = Session(eng, stack_cache=cache)
    srv.bind(path)
    srv.listen(128)
    srv.settimeout(5.0)
    sys.stderr.write(f"cspb200 server: GPU {device} listening on {path}\n")
    sys.stderr.flush()
    last = time.time()
    served = 0
    try:
        while True:
            try:
                conn, _ = srv.accept()
            except socket.timeout:
                if idle_timeout > 0 and time.time() - last > idle_timeout:
                    break
                continue
            with conn:
                f = conn.makefile("rwb")
                line = f.readline()
                if not line:
                    continue
                req = json.loads(line)
                if req.get("prog") == "shutdown":
                    f.write(json.dumps({"rc": 0, "out": f"served {served} requests\n", "err": ""}).encode() + b"\n")
                    f.flush()
                    break
                out, errbuf, err, rc = io.StringIO(), io.StringIO(), "", 0
                cwd0 = os.getcwd()
                env0 = {k: os.environ.get(k) for k in req.get("env", {})}
                t0 = time.time()
                try:
                    os.chdir(req.get("cwd") or cwd0)
                    os.environ.update(req.get("env", {}))
                    with contextlib.redirect_stderr(errbuf):
                        rc = _dispatch(req["prog"], list(req.get("argv", [])), req.get("stdin", ""), session, out) or 0
                except Exception as e:  # the word pyp greps for (particle_cspt.py:812-818)
                    rc = 1
                    err = f"{req.get('prog')}: caught error: {type(e).__name__}: {e}\n"
                    if os.environ.get("CSPB_SERVER_TRACE"):
                        err += traceback.format_exc()
                finally:
                    os.chdir(cwd0)
                    for k, v in env0.items():
                        if v is None:
                            os.environ.pop(k, None)
                        else:
                            os.environ[k] = v
                served += 1
                info = {"served": served, "seconds": time.time() - t0}
                if cache is not None:
                    info.update(stack_uploaded_mb=cache.uploaded_bytes / 1e6, stack_cache_hit_mb=cache.hit_bytes / 1e6)
                f.write(json.dumps({"rc": rc, "out": out.getvalue(), "err": errbuf.getvalue() + err, "server": info}).encode() + b"\n")
                f.flush()
                last = time.time()
    finally:
        srv.close()
        try:
            os.unlink(path)
        except OSError:
            pass
        eng.close()
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m pyp_b200.server")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--idle", type=float, default=float(os.environ.get("CSPB_SERVER_IDLE", 900)), help="exit after this many idle seconds (0 = never)")
    ap.add_argument("--stack-cache-gb", type=float, default=float(os.environ.get("CSPB_STACK_CACHE_GB", 64)))
    a = ap.parse_args(argv)
    return serve(a.device, a.idle, a.stack_cache_gb)


if __name__ == "__main__":
    sys.exit(main())
