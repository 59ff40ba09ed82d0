"""Import shim: the product package lives in ``py-tdgl_b200/`` (a name Python cannot
import directly); this package re-exports it as ``tdgl_b200``."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "py-tdgl_b200"))
from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: F401,E402
