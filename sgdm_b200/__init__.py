"""Importable alias of the `self-guided-diffusion-models_b200/` package.

The package directory carries the upstream project's name (with hyphens, which
Python cannot import), so this shim redirects `import sgdm_b200.*` into it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "self-guided-diffusion-models_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
