"""Import alias: the package directory is named ``llama-fft_b200`` (not a legal Python identifier), so this
shim makes it importable as ``llama_fft_b200`` by pointing ``__path__`` at it and running its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "llama-fft_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
