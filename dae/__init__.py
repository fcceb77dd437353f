"""`dae` — importable alias of the package directory ``dynamic-asr-eval_b200/``.

The directory name required by the repo layout contains a hyphen, which Python cannot
import.  This shim makes ``import dae`` / ``import dae.lib`` resolve into it.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dynamic-asr-eval_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
