"""Test-infrastructure stub of the un-vendored `bitorch` package (PyPI, unpinned in
/root/reference/requirements.txt:1).  It exists ONLY so that the reference's pure-Python n-bit path can be
imported in the build container by oracle/gen_golden.py; nothing in the product imports it.
Only the names the reference touches at import time are provided."""
from enum import Enum


class RuntimeMode(Enum):
    DEFAULT = 1
    CPU = 2
    GPU = 4
    INFERENCE_AUTO = 8

    def __add__(self, other):
        return self

    @staticmethod
    def available_values():
        return list(RuntimeMode)
