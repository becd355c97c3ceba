"""append_stacks drop-in (external/cistem2/append_stacks as pyp drives it from
src/pyp/inout/image/mrc.py:643-696 `merge_fast`): two stdin answers — the stack that grows, the stack that
is appended — and the word `Error:` in the output when the dimensions differ (the only thing the caller
checks).  Pure I/O: the slices are streamed onto the end of the first file and nz / mz are rewritten; this
is how the per-range stacks of `csp` mode -2 become the merged particle stack.
"""
import sys

import numpy as np

from ..formats import mrc
from .prompts import Answers, PromptError, banner


def parse(ans: Answers):
    return {"first": ans.text("input image file name #1"), "second": ans.text("input image file name #2")}


def run(p, out=sys.stdout, chunk=256):
    h1, h2 = mrc.read_header(p["first"]), mrc.read_header(p["second"])
    if (h1["nx"], h1["ny"]) != (h2["nx"], h2["ny"]):
        raise ValueError(f"{p['first']} ({h1['nx']}x{h1['ny']}) and {p['second']} ({h2['nx']}x{h2['ny']}) have different dimensions")
    out.write(banner("AppendStacks"))
    out.write("\nAdding Images...\n")
    for s in range(0, h2["nz"], chunk):
        e = min(h2["nz"], s + chunk)
        _, data = mrc.read(p["second"], first=s + 1, last=e)
        mrc.append(p["first"], np.ascontiguousarray(data, dtype=np.float32))
    out.write(f"\n{h2['nz']} images appended to {p['first']} ({h1['nz'] + h2['nz']} in total)\n")
    out.write("\nAppendStacks: Normal termination\n")


def main(argv=None):
    try:
        run(parse(Answers(program="append_stacks")))
    except (PromptError, ValueError, OSError) as e:
        # mrc.merge_fast raises when it finds "Error:" in the captured output (mrc.py:685-686)
        sys.stdout.write(f"Error: {e}\n")
        sys.stderr.write(f"append_stacks: caught error: {e}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
