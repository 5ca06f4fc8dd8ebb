"""g3py_b200 — B200-native exact-GP hot path behind g3py's operator surface.

    import g3py_b200 as g3
    gp = g3.GP(x, g3.Bias(), g3.SE(x))
    gp.observed(x, y)
    gp.logp(); gp.dlogp(); gp.find_MAP(); gp.predict()

The compute path is libg3b.so (hand-written CUDA for sm_100a, C ABI in include/g3b.h) called through
ctypes; there is no CPU fallback.
"""
from . import _cabi
from ._cabi import Context, G3Error, load as load_library, lib_path
from .hypers import Hypers, HyperVar, Registry, Freedom
from .hypers.kernels import *    # noqa: F401,F403
from .hypers.means import *      # noqa: F401,F403
from .hypers.mappings import *   # noqa: F401,F403
from .hypers.transports import *  # noqa: F401,F403
from .processes import *         # noqa: F401,F403
from .transport import *         # noqa: F401,F403

__version__ = "0.1.0"
