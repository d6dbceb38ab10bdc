"""Drop-in module: put this directory first on sys.path and the reference's own
``import datasets`` (src/main.py:3-4, src/gnngp.py:4, notebooks) resolves to the B200 implementation."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from gnngp_b200.datasets import *  # noqa: F401,F403,E402
from gnngp_b200 import datasets as _impl  # noqa: E402

globals().update({k: getattr(_impl, k) for k in dir(_impl) if k.startswith("_") and not k.startswith("__")})
