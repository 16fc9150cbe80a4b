import logging


class _Logger:
    def __init__(self, name):
        self._log = logging.getLogger(name)
        self._once = set()

    def __getattr__(self, item):
        return getattr(self._log, item)

    def warning_once(self, msg):
        if msg not in self._once:
            self._once.add(msg)
            self._log.warning(msg)


_loggers = {}


def getLogger(name):
    if name not in _loggers:
        _loggers[name] = _Logger(name)
    return _loggers[name]
