"""Import shim: the product lives in `dune-gdt_b200/` (not an importable name); this package exposes it as
`dune_gdt_b200` by extending its module search path."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dune-gdt_b200"))

from . import capi, descriptors  # noqa: E402
from .api import *  # noqa: E402,F401,F403
from .api import ApplyOn, Context, Stencil  # noqa: E402,F401
