"""Object checksums that key the on-disk caches (reference: mdevaluate/checksum.py:41-96).

The nojump cache (``.{traj}.nojump.npz``, reader.py:205-241) and the autosave result files
(autosave.py:138-183) store ``checksum(...)`` of what they were computed from.  This module
reproduces the reference's digest rule -- a salted SHA-1 over a type-dependent byte encoding of
every argument, returned as a big-endian integer -- so that for the argument kinds that do not
depend on source text (str, bytes, numbers, arrays, modules, objects with ``__checksum__``,
``functools.partial`` over those) the numbers are IDENTICAL to the reference's
(tests/golden/checksum.npz).  Plain Python functions hash their comment-stripped source plus the
values of their free variables, as in the reference; our functions have different source than the
reference's, so such checksums are compatible in rule, not in value.
"""
import functools
import hashlib
import inspect
from types import FunctionType, ModuleType

import numpy as np

from .logging import logger

# Salt of every digest.  Change it to invalidate all cached files at once.
SALT = 42


def version(version_nr, calls=[]):
    """Decorator: give a function a checksum that depends on its name, ``version_nr`` and ``calls``
    instead of on its source (checksum.py:16-27)."""
    def decorator(func):
        fixed = checksum(func.__name__, version_nr, *calls)

        @functools.wraps(func)
        def wrapped(*args, **kwargs):
            return func(*args, **kwargs)

        # functools.wraps copies func.__dict__, so setting it on the wrapped function is enough
        wrapped.__checksum__ = lambda: fixed
        func.__checksum__ = wrapped.__checksum__
        return wrapped
    return decorator


def strip_comments(source):
    """Source text without comment lines and docstrings (checksum.py:30-38): a line that starts
    with ``#`` or a quote is dropped; a line starting with a triple quote toggles docstring mode."""
    kept = []
    inside = False
    for line in source.split('\n'):
        head = line.strip()
        if head.startswith(('#', '"', "'")) or inside:
            opens = head.startswith(('"""', "'''"))
            inside = (opens + inside) == 1
            continue
        kept.append(line + '\n')
    return ''.join(kept)


def _feed(digest, arg, _seen=None):
    if hasattr(arg, '__checksum__'):
        logger.debug('Checksum via __checksum__: %s', str(arg))
        digest.update(str(arg.__checksum__()).encode())
    elif isinstance(arg, bytes):
        digest.update(arg)
    elif isinstance(arg, str):
        digest.update(arg.encode())
    elif isinstance(arg, ModuleType):
        digest.update(arg.__name__.encode())
    elif isinstance(arg, FunctionType):
        digest.update(strip_comments(inspect.getsource(arg)).encode())
        free = inspect.getclosurevars(arg)
        # the reference recurses into every referenced global; functions that refer to each other would never end
        # there, so a function already on the way is fed once (identical digests for anything without such cycles)
        seen = (_seen or ()) + (id(arg),)
        for value in {**free.nonlocals, **free.globals}.values():
            if value is not arg and not (isinstance(value, FunctionType) and id(value) in seen):
                _feed(digest, value, seen)
    elif isinstance(arg, functools.partial):
        logger.debug('Checksum via partial for %s', str(arg))
        _feed(digest, arg.func)
        for value in arg.args:
            _feed(digest, value)
        for key in sorted(arg.keywords.keys()):
            digest.update(key.encode())
            _feed(digest, arg.keywords[key])
    elif isinstance(arg, np.ndarray):
        digest.update(arg.tobytes())
    else:
        logger.debug('Checksum via str for %s', str(arg))
        digest.update(str(arg).encode())


def checksum(*args, csum=None):
    """Checksum of any objects as one integer (checksum.py:41-96); see the module docstring for the
    encoding of each kind of argument.  ``csum`` continues an existing digest."""
    if csum is None:
        csum = hashlib.sha1()
        csum.update(str(SALT).encode())
    for arg in args:
        _feed(csum, arg)
    return int.from_bytes(csum.digest(), 'big')
