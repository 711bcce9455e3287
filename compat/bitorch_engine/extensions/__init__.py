"""`bitorch_engine.extensions.<name>`: the module identity the reference resolves with importlib
(utils/safe_import.py:75-112)."""
EXTENSION_PREFIX = "bitorch_engine.extensions."
