"""Drop-in alias: puts the B200 implementation under the reference's module paths so that callers written against
GreenBitAI/bitorch-engine (green-bit-llm: `from bitorch_engine.layers.qlinear.nbit.cuda import MPQLinearCuda`,
`from bitorch_engine.optim import DiodeMix`, ...) resolve unchanged.  Use it by putting `<repo>/compat` on PYTHONPATH
(it is deliberately NOT importable from the repo root, where tests import the real reference for golden vectors).
Only the low-bit Linear hot path is provided; other reference modules raise ImportError."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

import bitorch_engine_b200 as _impl  # noqa: E402

__version__ = _impl.__version__


def initialize():
    """bitorch registration hook of the reference (bitorch_engine/__init__.py:1-6); nothing to register here."""
    return None
