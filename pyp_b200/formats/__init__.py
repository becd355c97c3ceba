"""Wire formats of the hot path: `.cistem` parameter tables and MRC stacks/volumes."""
from . import cistem, dump, mrc, statistics  # noqa: F401
