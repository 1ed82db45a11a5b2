"""Package logger (same logger name as the reference, mdevaluate/logging.py:1-26)."""
import logging

logger = logging.getLogger('mdevaluate')
if not logger.handlers:
    _handler = logging.StreamHandler()
    _handler.setFormatter(logging.Formatter('[%(asctime)s] %(levelname)s:%(name)s: %(message)s'))
    logger.addHandler(_handler)
    logger.setLevel(logging.INFO)


def setlevel(level, file=None):
    """Change the level of the package logger, optionally logging to ``file`` too."""
    if isinstance(level, str):
        level = getattr(logging, level.upper())
    logger.setLevel(level)
    if file is not None:
        fh = logging.FileHandler(file)
        fh.setLevel(level)
        logger.addHandler(fh)
