"""Importable name of the package in `factor-graph-neural-network_b200/` (that directory name,
required by the repo layout, is not a Python identifier).  `import fgnn_b200` runs the real
package's __init__ with this module as its namespace; submodules resolve through __path__."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "factor-graph-neural-network_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _os, _real
