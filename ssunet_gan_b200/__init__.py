"""Importable alias of the ``ssunet-gan_b200/`` package directory (a hyphen is not a valid
Python identifier): ``import ssunet_gan_b200`` exposes everything under ssunet-gan_b200/."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ssunet-gan_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
