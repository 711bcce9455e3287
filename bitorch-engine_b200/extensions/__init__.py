"""Python-visible twins of the reference's extension modules (`bitorch_engine.extensions.<name>`,
bitorch_engine/extensions/__init__.py:1, loaded by utils/safe_import.py:75-112).  Same function names, argument
order and return types; each function validates, allocates the output with torch and calls the C ABI on the
current CUDA stream."""
EXTENSION_PREFIX = "bitorch_engine.extensions."
