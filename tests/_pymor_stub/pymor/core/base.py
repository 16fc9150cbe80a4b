"""BasicObject / ImmutableObject: `self.__auto_init(locals())`, `with_`, `logger`, `name`."""
import abc
import inspect
import itertools

from pymor.core.exceptions import ConstError
from pymor.core.logger import getLogger

abstractmethod = abc.abstractmethod
_uid = itertools.count()


class UberMeta(abc.ABCMeta):
    def __init__(cls, name, bases, namespace):
        super().__init__(name, bases, namespace)
        init = getattr(cls, "__init__", None)
        try:
            params = list(inspect.signature(init).parameters.values())[1:]
            args = tuple(p.name for p in params if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY))
        except (TypeError, ValueError):
            args = ()
        cls._init_arguments = args

        def __auto_init(self, locals_, _args=args):
            for arg in _args:
                if arg not in self.__dict__:
                    setattr(self, arg, locals_[arg])
        # `self.__auto_init` inside class X is name-mangled to `self._X__auto_init`
        setattr(cls, f"_{name.lstrip('_')}__auto_init", __auto_init)


class BasicObject(metaclass=UberMeta):
    _name = None

    @property
    def name(self):
        return self._name if self._name is not None else type(self).__name__

    @name.setter
    def name(self, value):
        self._name = value

    @property
    def logger(self):
        return getLogger(f"{type(self).__module__}.{type(self).__name__}")

    @property
    def uid(self):
        if "_uid" not in self.__dict__:
            self.__dict__["_uid"] = next(_uid)
        return self.__dict__["_uid"]


class ImmutableMeta(UberMeta):
    def __call__(cls, *args, **kwargs):
        obj = super().__call__(*args, **kwargs)
        obj.__dict__["_locked"] = True
        return obj


class ImmutableObject(BasicObject, metaclass=ImmutableMeta):
    """Attributes not starting with '_' are frozen once __init__ has returned."""
    _locked = False

    def __setattr__(self, key, value):
        if self._locked and not key.startswith("_"):
            raise ConstError(f"changing {key!r} of immutable {type(self).__name__}")
        object.__setattr__(self, key, value)

    def with_(self, new_type=None, **kwargs):
        c = type(self) if new_type is None else new_type
        unknown = set(kwargs) - set(c._init_arguments)
        if unknown:
            raise ConstError(f"with_: {unknown} are not __init__ arguments of {c.__name__}")
        for arg in c._init_arguments:
            if arg not in kwargs:
                kwargs[arg] = getattr(self, arg)
        return c(**kwargs)
