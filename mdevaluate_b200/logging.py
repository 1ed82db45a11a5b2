"""The package logger.

It shares the logger name ``'mdevaluate'`` with the reference so that applications which
configure that logger keep working; ``setlevel`` is the one helper the reference exposes.
"""
import logging as _stdlib_logging

LOGGER_NAME = 'mdevaluate'
_FORMAT = '[%(asctime)s] %(levelname)s:%(name)s: %(message)s'

logger = _stdlib_logging.getLogger(LOGGER_NAME)


def _attach(handler, level=None):
    handler.setFormatter(_stdlib_logging.Formatter(_FORMAT))
    if level is not None:
        handler.setLevel(level)
    logger.addHandler(handler)
    return handler


if not logger.handlers:          # importing twice (or next to the reference) must not duplicate output
    _attach(_stdlib_logging.StreamHandler())
    logger.setLevel(_stdlib_logging.INFO)


def setlevel(level, file=None):
    """Set the verbosity (a ``logging`` level or its name); with ``file`` also log into that file."""
    numeric = _stdlib_logging.getLevelName(level.upper()) if isinstance(level, str) else level
    logger.setLevel(numeric)
    if file is not None:
        _attach(_stdlib_logging.FileHandler(file), numeric)
