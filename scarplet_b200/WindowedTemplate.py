"""Drop-in module name: ``from scarplet_b200.WindowedTemplate import Scarp`` mirrors
``from scarplet.WindowedTemplate import Scarp``.  See ``templates.py``."""
from .templates import (Channel, LeftFacingUpperBreakScarp, Ricker,  # noqa: F401
                        RightFacingUpperBreakScarp, Scarp, WindowedTemplate)
