"""Answer-list parsing shared by the stdin-driven front-ends.  pyp feeds the cisTEM-style
binaries one answer per line through a heredoc (src/pyp/refine/frealign/frealign.py:3918-3994,
1780-1824, 1878-1888, 2075-2093); yes/no are the literals ``yes``/``no``, numbers are Python
``str()`` renderings."""
import os
import sys


class Answers:
    def __init__(self, text=None, program="program"):
        if text is None:
            text = sys.stdin.read()
        self.lines = [ln.strip() for ln in text.splitlines()]
        while self.lines and self.lines[-1] == "":
            self.lines.pop()
        self.pos = 0
        self.program = program

    def _next(self, what):
        if self.pos >= len(self.lines):
            raise PromptError(f"{self.program}: ran out of answers while reading '{what}' (answer #{self.pos + 1})")
        v = self.lines[self.pos]
        self.pos += 1
        return v

    def text(self, what):
        return self._next(what)

    def number(self, what):
        v = self._next(what)
        try:
            return float(v)
        except ValueError:
            raise PromptError(f"{self.program}: answer #{self.pos} for '{what}' is not a number: {v!r}")

    def integer(self, what):
        return int(round(self.number(what)))

    def yesno(self, what):
        v = self._next(what).lower()
        if v in ("yes", "y", "true", "t", "1"):
            return True
        if v in ("no", "n", "false", "f", "0"):
            return False
        raise PromptError(f"{self.program}: answer #{self.pos} for '{what}' must be yes/no, got {v!r}")

    def done(self):
        return self.pos >= len(self.lines)


class PromptError(ValueError):
    pass


def pick_device(first=1, count=1):
    """GPU for this invocation.  pyp forks one process per contiguous range
    (src/pyp/system/local_run.py:507-516); ranges are spread round-robin over the visible GPUs
    unless CSPB_DEVICE pins one."""
    if "CSPB_DEVICE" in os.environ:
        return int(os.environ["CSPB_DEVICE"])
    n = int(os.environ.get("CSPB_NUM_DEVICES", "0"))
    if n <= 0:
        try:
            from .._lib import lib  # the engine's own count: no framework import in the front-ends

            n = int(lib().cspb_device_count())
        except Exception:
            n = 1
    n = max(1, n)
    return ((int(first) - 1) // max(1, int(count))) % n


def write_notes(out, program, items):
    """One log line per answer that was read but does not change the computation (DESIGN.md §7,
    oracle/SEMANTICS.md §8): `items` = [(condition, text), ...].  Returns the number of notes written."""
    n = 0
    for cond, text in items:
        if cond:
            out.write(f"{program}: note: {text}\n")
            n += 1
    return n


def banner(name):
    return (f"\n        **   Welcome to {name}   **\n\n"
            "            Version : cspb200 0.1.0 (B200-native drop-in)\n"
            "   Library Version : libcspb200 ABI 1\n"
            "               Mode : Scripted\n")
