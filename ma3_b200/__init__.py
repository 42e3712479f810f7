"""Import alias: the product package lives in `make-an-audio-3_b200/` (not a valid Python identifier), so YAML
`target:` strings and tests import it as `ma3_b200.*`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "make-an-audio-3_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
