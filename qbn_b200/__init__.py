"""Import alias: the product package lives in `quantised-bayesian-nets_b200/` (a directory name
Python cannot import directly because of the hyphens); `import qbn_b200` loads it from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "quantised-bayesian-nets_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
