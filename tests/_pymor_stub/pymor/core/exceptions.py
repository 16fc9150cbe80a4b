class ConstError(Exception):
    pass


class AccuracyError(Exception):
    pass


class RuleNotMatchingError(NotImplementedError):
    pass


class NoMatchingRuleError(NotImplementedError):
    def __init__(self, obj):
        super().__init__(f"no rule matches {type(obj).__name__}")
        self.obj = obj


class InversionError(Exception):
    pass
