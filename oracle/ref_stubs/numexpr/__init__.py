"""Import stand-in for numexpr (see ../README.md)."""
import sys
import numpy as _np

_FUNCS = {name: getattr(_np, name) for name in
          ("real", "imag", "abs", "cos", "sin", "exp", "sqrt", "where", "log")}


def evaluate(expr, local_dict=None, global_dict=None):
    frame = sys._getframe(1)
    scope = dict(frame.f_globals if global_dict is None else global_dict)
    scope.update(_FUNCS)
    scope.update(frame.f_locals if local_dict is None else local_dict)
    return eval(" ".join(expr.split()), {"__builtins__": {}}, scope)
