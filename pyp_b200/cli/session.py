"""What a front-end needs from the process it runs in: an Engine and the images of a stack range.

pyp spawns one short-lived process per particle range (src/pyp/system/local_run.py:507-516); each of our
stand-alone front-ends therefore creates a throw-away :class:`Session` (new CUDA context, reference transformed
again, stack range read from the file).  The resident server (pyp_b200/server.py) keeps ONE session per GPU
alive between invocations: the context, the transformed reference (`Engine.ensure_reference`) and the stack
ranges already uploaded to HBM are reused, so a front-end call costs its own kernels and little else."""
import numpy as np

from ..formats import mrc
from .prompts import pick_device


class Session:
    def __init__(self, eng=None, stack_cache=None):
        self._eng = eng
        self._owned = eng is None
        self.stack_cache = stack_cache  # callable(path, positions_1based) -> images (numpy or CUDA tensor), or None

    def engine(self, first=1, count=1):
        if self._eng is None:
            from ..engine import Engine

            self._eng = Engine(pick_device(first, count))
        return self._eng

    def images(self, path, positions):
        """Images at the 1-based stack positions (ascending), as something Engine.load_images / recon_insert accept."""
        pos = np.asarray(positions, dtype=np.int64)
        if self.stack_cache is not None:
            return self.stack_cache(path, pos)
        _, data = mrc.read(path, first=int(pos.min()), last=int(pos.max()))
        return np.ascontiguousarray(data[pos - pos.min()])

    def release(self):
        if self._owned and self._eng is not None:
            self._eng.close()
            self._eng = None
