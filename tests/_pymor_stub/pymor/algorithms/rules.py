"""Rule tables: ordered lists of (matcher, action); `insert_rule(index, rule)` as the reference
uses it (rla/__init__.py:20-21, utilities/__init__.py:71-72)."""
from pymor.core.base import BasicObject, UberMeta
from pymor.core.exceptions import NoMatchingRuleError, RuleNotMatchingError
from pymor.operators.interface import Operator


class rule:
    def __call__(self, action):
        self.action = action
        self.name = action.__name__
        return self

    def matches(self, obj):
        raise NotImplementedError


class match_class(rule):
    def __init__(self, *classes):
        self.classes = classes

    def matches(self, obj):
        return isinstance(obj, self.classes)


class match_generic(rule):
    def __init__(self, condition, condition_description=None):
        self.condition = condition

    def matches(self, obj):
        return bool(self.condition(obj))


class match_always(rule):
    def __init__(self, action=None):
        if action is not None:
            self(action)

    def matches(self, obj):
        return True


class RuleTableMeta(UberMeta):
    def __new__(mcs, name, bases, namespace):
        rules = [v for v in namespace.values() if isinstance(v, rule)]      # definition order
        namespace["rules"] = rules
        return super().__new__(mcs, name, bases, namespace)


class RuleTable(BasicObject, metaclass=RuleTableMeta):
    def __init__(self, use_caching=False):
        self._cache = {} if use_caching else None

    @classmethod
    def insert_rule(cls, index, rule_):
        assert isinstance(rule_, rule)
        cls.rules.insert(index, rule_)

    @classmethod
    def append_rule(cls, rule_):
        cls.rules.append(rule_)

    def apply(self, obj):
        for r in self.rules:
            if r.matches(obj):
                try:
                    return r.action(self, obj)
                except RuleNotMatchingError:
                    continue
        raise NoMatchingRuleError(obj)

    @staticmethod
    def get_children(obj):
        children = []
        for arg in obj._init_arguments:
            v = getattr(obj, arg, None)
            if isinstance(v, Operator):
                children.append(arg)
            elif isinstance(v, (list, tuple)) and len(v) and all(isinstance(o, Operator) for o in v):
                children.append(arg)
        return children

    def apply_children(self, obj, children=None):
        children = self.get_children(obj) if children is None else children
        out = {}
        for name in children:
            v = getattr(obj, name)
            out[name] = type(v)(self.apply(o) for o in v) if isinstance(v, (list, tuple)) else self.apply(v)
        return out

    def replace_children(self, obj, children=None):
        new = self.apply_children(obj, children)
        if all(getattr(obj, k) is v or (isinstance(v, (list, tuple)) and all(a is b for a, b in zip(getattr(obj, k), v)))
               for k, v in new.items()):
            return obj
        return obj.with_(**new)
