"""Import alias: the product package lives in the directory `bitorch-engine_b200/` (not a valid Python identifier);
`import bitorch_engine_b200` resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bitorch-engine_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
