"""Importable alias of the package directory ``warm-start-hybrid-mpc_b200/`` (a hyphenated
directory name cannot be imported directly): ``import warm_start_hmpc_b200 as ws``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'warm-start-hybrid-mpc_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f
