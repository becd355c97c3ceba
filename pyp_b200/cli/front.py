"""Thin client of the resident engine (pyp_b200/server.py) — standard library only, started with `python -S` by the
executables in bin/ when CSPB_SERVER is set, so that an invocation costs ~40 ms of interpreter start-up instead of
~0.5 s of numpy import + 1.4 s of CUDA context.

CSPB_SERVER=1     use the daemon of the chosen GPU if it is listening, else run stand-alone
CSPB_SERVER=auto  as above, but start the daemon first when it is not there (it stays up for later invocations)

The particle range (answers `first` / `last`) picks the GPU exactly like the stand-alone front-ends do
(prompts.pick_device): ranges are spread round-robin over the visible GPUs unless CSPB_DEVICE pins one.
"""
import glob
import json
import os
import socket
import subprocess
import sys
import tempfile
import time

# (index of `first`, index of `last`) in the stdin answer lists — frealign.py:3918-3994, 1780-1824, 3995-4041
_RANGE_AT = {"refine3d": (11, 12), "reconstruct3d": (9, 10)}
_STDIN_PROGS = ("refine3d", "reconstruct3d", "merge3d", "local_merge3d", "refine_ctf")
_PASS_ENV = ("PYP_SCRATCH", "CSPB_DEVICE", "CSPB_NUM_DEVICES")


def socket_path(device):
    d = os.environ.get("CSPB_SOCKET_DIR") or tempfile.gettempdir()
    return os.path.join(d, f"cspb200-{os.getuid()}-gpu{int(device)}.sock")


def n_devices():
    n = int(os.environ.get("CSPB_NUM_DEVICES", "0") or 0)
    if n > 0:
        return n
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis is not None and vis.strip() != "":
        return max(1, len([v for v in vis.split(",") if v.strip()]))
    return max(1, len(glob.glob("/dev/nvidia[0-9]*")))


def pick_device(prog, text, argv):
    if "CSPB_DEVICE" in os.environ:
        return int(os.environ["CSPB_DEVICE"])
    first, count = 1, 1
    try:
        if prog in _RANGE_AT:
            lines = [ln.strip() for ln in text.splitlines()]
            a, b = _RANGE_AT[prog]
            first, last = int(float(lines[a])), int(float(lines[b]))
            count = max(1, last - first + 1)
        elif prog in ("csp", "csp_GS") and len(argv) >= 5:
            first, last = int(argv[3]) + 1, int(argv[4]) + 1
            count = max(1, last - first + 1)
    except (ValueError, IndexError):
        pass
    return ((first - 1) // count) % n_devices()


def _connect(path, timeout):
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(timeout)
    s.connect(path)
    return s


def _spawn_server(device):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    log = os.path.join(os.environ.get("CSPB_SOCKET_DIR") or tempfile.gettempdir(), f"cspb200-{os.getuid()}-gpu{device}.log")
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.pop("CSPB_SERVER", None)
    python = os.environ.get("CSPB_PYTHON") or sys.executable
    with open(log, "ab") as lf:
        subprocess.Popen([python, "-m", "pyp_b200.server", "--device", str(device)], stdin=subprocess.DEVNULL, stdout=lf, stderr=lf,
                         env=env, start_new_session=True, cwd=root)


def _run_local(prog, argv):
    # stand-alone front-end in a full interpreter (this one runs with -S: no site-packages); stdin is still unread
    python = os.environ.get("CSPB_PYTHON") or sys.executable
    mod = "csp" if prog == "csp_GS" else prog
    os.execv(python, [python, "-m", f"pyp_b200.cli.{mod}"] + list(argv))


def main():
    if len(sys.argv) < 2:
        sys.stderr.write("usage: python -S -m pyp_b200.cli.front <program> [args]\n")
        return 2
    prog, argv = sys.argv[1], sys.argv[2:]
    mode = os.environ.get("CSPB_SERVER", "").lower()
    if mode in ("", "0", "no", "off"):
        _run_local(prog, argv)
    text = sys.stdin.read() if prog in _STDIN_PROGS else ""
    device = pick_device(prog, text, argv)
    path = socket_path(device)
    sock = None
    try:
        sock = _connect(path, 5.0)
    except OSError:
        if mode == "auto":
            # several first callers may race: only one daemon binds the socket, the others fail on bind and exit
            _spawn_server(device)
            t_end = time.time() + float(os.environ.get("CSPB_SERVER_START_TIMEOUT", 120))
            while time.time() < t_end:
                try:
                    sock = _connect(path, 5.0)
                    break
                except OSError:
                    time.sleep(0.2)
    if sock is None:
        # no daemon: the stand-alone path needs the answers on stdin again
        python = os.environ.get("CSPB_PYTHON") or sys.executable
        mod = "csp" if prog == "csp_GS" else prog
        p = subprocess.run([python, "-m", f"pyp_b200.cli.{mod}"] + list(argv), input=text.encode())
        return p.returncode
    sock.settimeout(float(os.environ.get("CSPB_SERVER_TIMEOUT", 86400)))
    req = {"prog": prog, "argv": argv, "cwd": os.getcwd(), "stdin": text, "env": {k: os.environ[k] for k in _PASS_ENV if k in os.environ}}
    f = sock.makefile("rwb")
    f.write(json.dumps(req).encode() + b"\n")
    f.flush()
    line = f.readline()
    sock.close()
    if not line:
        sys.stderr.write(f"{prog}: caught error: the cspb200 server closed the connection\n")
        return 1
    resp = json.loads(line)
    sys.stdout.write(resp.get("out", ""))
    sys.stderr.write(resp.get("err", ""))
    return int(resp.get("rc", 1))


if __name__ == "__main__":
    sys.exit(main())
