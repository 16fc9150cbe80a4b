import numpy as np


class Mu(dict):
    """Parameter values: name -> 1-D array."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            dict.__setitem__(self, k, np.atleast_1d(np.asarray(v, dtype=float)))

    def __hash__(self):
        return hash(tuple((k, tuple(v)) for k, v in sorted(self.items())))


class ParametricObject:
    """`parametric` is derived from the __init__ arguments that are themselves parametric."""
    _own_parametric = False

    @property
    def parametric(self):
        if self._own_parametric:
            return True
        for arg in getattr(self, "_init_arguments", ()):
            v = self.__dict__.get(arg)
            items = v if isinstance(v, (list, tuple)) else (v,)
            for it in items:
                if isinstance(it, ParametricObject) and it is not self and it.parametric:
                    return True
        return False
