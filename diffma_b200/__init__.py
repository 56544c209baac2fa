"""Importable alias of the product package.

The product lives in ``diffma-diffusion-mamba_b200/`` (a directory name Python cannot import
because of the hyphens); this stub makes it importable as ``diffma_b200`` by pointing the
package search path at that directory and running its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "diffma-diffusion-mamba_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
