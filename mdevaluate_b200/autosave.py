"""Result files for expensive evaluations -- the reference's ``autosave`` workflow (mdevaluate/autosave.py,
doc/special/autosave.rst) on top of a small ``ResultFile`` abstraction.

``enable(dir)`` switches it on; from then on a function wrapped by ``autosave_data`` --
``correlation.shifted_correlation`` is -- stores ``(times, result)`` in

    <trajectory dir>/<dir>/<function>_<correlation function>_<description>.npz

next to a checksum of the arguments that determine the result, and a later call with the same
arguments returns the stored arrays instead of launching anything.  What is kept compatible with the
reference (autosave.py:85-183) is what its users and tooling see: the public names of this module,
the file names, the npz layout (entries ``checksum`` and ``data``) and the digest rule (checksum.py).
Checksums of plain Python functions hash their source text and therefore differ between the two
packages: a result the reference stored under such a key is recomputed here, never silently trusted.
"""
import functools
import inspect
import os

import numpy as np

from .checksum import checksum
from .logging import logger

# Module state, readable like the reference's globals (autosave.py:8-11); enable() / disable() set it.
autosave_directory = None
load_autosave_data = False
verbose_print = True
user_autosave_directory = os.path.join(os.environ.get('HOME', os.path.expanduser('~')), '.mdevaluate/autosave')


def notify(msg):
    """Progress messages: INFO when ``enable(..., verbose=True)``, else DEBUG."""
    logger.log(20 if verbose_print else 10, msg)


def enable(dir, load_data=True, verbose=True):
    """Store results of decorated functions below ``dir`` -- a path relative to each trajectory's own
    directory -- and, unless ``load_data=False``, reuse stored results whose checksum matches."""
    global autosave_directory, load_autosave_data, verbose_print
    autosave_directory, load_autosave_data, verbose_print = dir, load_data, verbose
    notify('Enabled autosave in directory: {}'.format(autosave_directory))


def disable():
    """No result files are written or read until the next ``enable``."""
    global autosave_directory, load_autosave_data
    autosave_directory, load_autosave_data = None, False


class disabled:
    """``with autosave.disabled(): ...`` -- autosave is off inside the block and back on, with the
    directory it had, afterwards."""

    def __enter__(self):
        self._autosave_directory = autosave_directory
        disable()
        return self

    def __exit__(self, *exc):
        enable(self._autosave_directory)
        return False


trust_pickle = False
"""Result files written by the REFERENCE hold their checksum (a 160-bit Python integer) and ragged results only as
pickled object entries.  Unpickling runs code from the file, so such files are read only after an explicit
``autosave.trust_pickle = True``; files written by this package carry pickle-free twins of every entry
(``checksum_str``, ``item_<k>``) and are always read with ``allow_pickle=False``."""


class ResultFile:
    """One ``.npz`` result file.  Layout of the reference (autosave.py:85-100): ``checksum`` and ``data``; next to
    them this package writes ``checksum_str`` (the checksum in decimal) and, for a ragged result, ``n_items`` and
    ``item_<k>`` -- the entries it reads back, none of which needs pickle."""

    def __init__(self, path):
        self.path = path

    def stored_checksum(self):
        """The checksum inside the file; 0 when the file is absent, carries none, or carries only a pickled one
        and ``trust_pickle`` is off."""
        if not os.path.exists(self.path):
            return 0
        with np.load(self.path, allow_pickle=False) as archive:
            if 'checksum_str' in archive:
                return int(str(archive['checksum_str'][()]))
            if 'checksum' not in archive:
                return 0
        if not trust_pickle:
            logger.info('ignoring %s: written without checksum_str (set autosave.trust_pickle = True to read files '
                        'of the reference)', self.path)
            return 0
        with np.load(self.path, allow_pickle=True) as archive:
            return archive['checksum'][()]

    def holds(self, *checksums):
        stored = self.stored_checksum()
        return any(stored == c for c in checksums)

    def write(self, csum, result):
        """A tuple of equally shaped arrays becomes one array; anything ragged -- ``(times (L,), result
        (S, L))``, the ``(times, result, count)`` of multi-sub-ensemble runs -- an object array, one
        entry per item (the reference's layout), plus the pickle-free ``item_<k>`` entries."""
        notify('Saving result to file: {}'.format(self.path))
        extra = {'checksum_str': np.array(str(int(csum)))}
        try:
            packed = np.array(result)
            if packed.dtype == object:
                raise ValueError
        except ValueError:
            packed = np.empty((len(result),), dtype=object)
            for slot, item in enumerate(result):
                packed[slot] = item
                extra['item_{}'.format(slot)] = np.asarray(item)
            extra['n_items'] = np.array(len(result))
        np.savez(self.path, checksum=csum, data=packed, **extra)

    def read(self):
        """The stored result.  Files from before the ``data`` entry existed hold positional ``arr_*``
        entries instead; they are returned as a tuple and rewritten in today's layout."""
        notify('Loading result from file: {}'.format(self.path))
        with np.load(self.path, allow_pickle=False) as archive:
            if 'n_items' in archive:
                items = np.empty((int(archive['n_items']),), dtype=object)
                for slot in range(len(items)):
                    items[slot] = archive['item_{}'.format(slot)]
                return items
            if 'checksum_str' in archive and 'data' in archive:
                return archive['data']
        if not trust_pickle:
            raise ValueError('{} holds pickled entries only; set autosave.trust_pickle = True to read result files '
                             'written by the reference'.format(self.path))
        with np.load(self.path, allow_pickle=True) as archive:
            if 'data' in archive:
                return archive['data']
            legacy = tuple(archive[name] for name in sorted(archive) if 'arr' in name)
            csum = archive['checksum'][()]
        self.write(csum, legacy)
        return legacy


def get_directory(reader):
    """Where the results of ``reader``'s trajectory go: ``<dirname(reader.filename)>/<autosave dir>``,
    created on demand; a read-only location is mirrored below ``~/.mdevaluate/autosave``."""
    target = os.path.join(os.path.dirname(reader.filename), autosave_directory)
    try:
        os.makedirs(target, exist_ok=True)
    except PermissionError:
        pass
    if not os.access(target, os.W_OK):
        target = os.path.join(user_autosave_directory, target.lstrip('/'))
        logger.info('Switched autosave directory to {}, since original location is not writeable.'.format(target))
        os.makedirs(target, exist_ok=True)
    return target


def get_filename(function, checksum, description, *args):
    """File of one call.  The stem names the function and every argument that has a name (a function, or
    the function inside a ``partial``); the description is the caller's plus the ``.description`` of the
    arguments (the atom selection of a Coordinates); the directory comes from the argument that carries a
    reader in ``.frames``.  E.g. ``shifted_correlation_isf_OW.npz``."""
    stem = [function.__name__]
    notes = [description]
    directory = None
    for arg in args:
        name = getattr(arg, '__name__', None)
        if name is None and isinstance(arg, functools.partial):
            name = arg.func.__name__
        if name is not None:
            stem.append(name)
        if hasattr(arg, 'frames'):
            directory = get_directory(arg.frames)
        note = getattr(arg, 'description', '')
        if note != '':
            notes.append(note)
    if directory is None:
        raise ValueError('autosave needs an argument with a .frames reader (a Coordinates object) to place the file')
    return os.path.join(directory, '{}_{}.npz'.format('_'.join(stem).strip('_'), '_'.join(notes).strip('_')))


def verify_file(filename, checksum):
    """Whether ``filename`` holds the result of the call with this checksum."""
    return ResultFile(filename).holds(checksum)


def save_data(filename, checksum, data):
    """Write ``checksum`` and ``data`` to ``filename`` (see :meth:`ResultFile.write`)."""
    ResultFile(filename).write(checksum, data)


def load_data(filename):
    """Read the result stored in ``filename`` (see :meth:`ResultFile.read`)."""
    return ResultFile(filename).read()


def autosave_data(nargs, kwargs_keys=None, version=None):
    """Decorator factory: keep the decorated function's results in result files while autosave is enabled.

    Args:
        nargs: how many leading positional arguments determine the result
        kwargs_keys (opt.): keyword arguments that determine the result as well
        version (opt.): a tag that stands in for the function's source in the checksum, so that stored
            results survive changes of the implementation that do not change the numbers

    The wrapper takes two keywords of its own: ``description`` (goes into the file name) and
    ``autoload=False`` (recompute even if a matching file exists).
    """
    def decorate(function):
        leading = list(inspect.signature(function).parameters)[:nargs]      # may also arrive as keywords
        watched = leading + list(kwargs_keys or ()) if kwargs_keys is not None else []

        @functools.wraps(function)
        def wrapper(*args, **kwargs):
            description = kwargs.pop('description', '') or ''
            reuse = kwargs.pop('autoload', True) and load_autosave_data
            if autosave_directory is None:
                return function(*args, **kwargs)
            relevant = list(args[:nargs]) + [kwargs[key] for key in watched if key in kwargs]
            by_source = checksum(function, *relevant)
            key = by_source if version is None else checksum(version, *relevant)
            target = ResultFile(get_filename(function, key, description, *relevant))
            if reuse and target.holds(key, by_source):
                return target.read()
            result = function(*args, **kwargs)
            target.write(key, result)
            return result
        return wrapper
    return decorate
