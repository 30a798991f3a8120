"""Alias of ``models.pl`` (the reference README spells the directory ``PL``, README.md:73; on disk it is ``pl``)."""
import sys as _sys

from .. import pl as _pl
from ..pl import models  # noqa: F401

_sys.modules[__name__ + '.models'] = _pl.models
