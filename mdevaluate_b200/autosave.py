"""Result files for expensive evaluations (reference: mdevaluate/autosave.py).

``enable(dir)`` switches it on; from then on a function decorated with ``autosave_data`` --
``correlation.shifted_correlation`` is -- writes ``(times, result)`` to
``<trajectory dir>/<dir>/<function>_<correlation function>_<description>.npz`` together with a
checksum of its relevant arguments, and returns the stored data instead of recomputing when a file
with the same checksum exists.  File names, the npz layout (``checksum``, ``data``) and the checksum
rule (checksum.py) are the reference's (autosave.py:85-183), so its tooling keeps working on the
files; checksums of plain Python functions depend on their source text and therefore differ between
the two packages -- a result the reference stored is recomputed here, not silently trusted.
"""
import functools
import inspect
import os

import numpy as np

from .checksum import checksum
from .logging import logger

autosave_directory = None
load_autosave_data = False
verbose_print = True
user_autosave_directory = os.path.join(os.environ.get('HOME', os.path.expanduser('~')), '.mdevaluate/autosave')


def notify(msg):
    (logger.info if verbose_print else logger.debug)(msg)


def enable(dir, load_data=True, verbose=True):
    """Save the results of decorated functions below ``dir`` (relative to the trajectory's directory)
    and, with ``load_data``, load matching results instead of recomputing (autosave.py:24-39)."""
    global autosave_directory, load_autosave_data, verbose_print
    verbose_print = verbose
    autosave_directory = dir
    load_autosave_data = load_data
    notify('Enabled autosave in directory: {}'.format(autosave_directory))


def disable():
    """Switch autosave off (autosave.py:42-46)."""
    global autosave_directory, load_autosave_data
    autosave_directory = None
    load_autosave_data = False


class disabled:
    """Context manager: no autosave inside the block; the previous directory is re-enabled afterwards
    (autosave.py:49-70)."""

    def __enter__(self):
        self._autosave_directory = autosave_directory
        disable()

    def __exit__(self, *args):
        enable(self._autosave_directory)


def get_directory(reader):
    """The autosave directory for a trajectory reader, created on demand; falls back to
    ``~/.mdevaluate/autosave/...`` when the trajectory's directory is read-only (autosave.py:73-86)."""
    savedir = os.path.join(os.path.dirname(reader.filename), autosave_directory)
    if not os.path.exists(savedir):
        try:
            os.makedirs(savedir)
        except PermissionError:
            pass
    if not os.access(savedir, os.W_OK):
        savedir = os.path.join(user_autosave_directory, savedir.lstrip('/'))
        logger.info('Switched autosave directory to {}, since original location is not writeable.'.format(savedir))
    os.makedirs(savedir, exist_ok=True)
    return savedir


def get_filename(function, checksum, description, *args):
    """``<savedir>/<function>[_<name of each named argument>]_<description>[_<argument descriptions>].npz``
    (autosave.py:89-104); the directory comes from the argument that has ``.frames`` (the Coordinates)."""
    func_desc = function.__name__
    savedir = None
    for arg in args:
        if hasattr(arg, '__name__'):
            func_desc += '_{}'.format(arg.__name__)
        elif isinstance(arg, functools.partial):
            func_desc += '_{}'.format(arg.func.__name__)
        if hasattr(arg, 'frames'):
            savedir = get_directory(arg.frames)
        if hasattr(arg, 'description') and arg.description != '':
            description += '_{}'.format(arg.description)
    if savedir is None:
        raise ValueError('autosave needs an argument with a .frames reader (a Coordinates object) to place the file')
    filename = '{}_{}.npz'.format(func_desc.strip('_'), description.strip('_'))
    return os.path.join(savedir, filename)


def _stored_checksum(filename):
    if not os.path.exists(filename):
        return 0
    # the checksum is a 160-bit Python integer, which NumPy stores as a pickled object entry
    with np.load(filename, allow_pickle=True) as data:
        return data['checksum'][()] if 'checksum' in data else 0


def verify_file(filename, checksum):
    """Whether ``filename`` holds the result of the call with this checksum (autosave.py:107-114)."""
    return _stored_checksum(filename) == checksum


def save_data(filename, checksum, data):
    """Write ``checksum`` and ``data`` to an npz file (autosave.py:117-127); a ragged result (e.g. the
    ``(times, result, count)`` of multi-subensemble runs) is stored as an object array."""
    notify('Saving result to file: {}'.format(filename))
    try:
        data = np.array(data)
    except ValueError:
        arr = np.empty((len(data),), dtype=object)
        arr[:] = data
        data = arr
    np.savez(filename, checksum=checksum, data=data)


def load_data(filename):
    """Read the result back (autosave.py:130-139); old files with positional ``arr_*`` entries are
    rewritten in the current layout."""
    notify('Loading result from file: {}'.format(filename))
    with np.load(filename, allow_pickle=True) as fdata:
        if 'data' in fdata:
            return fdata['data']
        data = tuple(fdata[k] for k in sorted(fdata) if ('arr' in k))
        stored = fdata['checksum'][()]
    save_data(filename, stored, data)
    return data


def autosave_data(nargs, kwargs_keys=None, version=None):
    """Decorator: autosave the result of a function (autosave.py:142-183).

    Args:
        nargs: number of leading positional arguments that determine the result
        kwargs_keys (opt.): keyword arguments that determine the result
        version (opt.): replaces the checksum of the function's code, so the files stay valid when
            the implementation changes without changing results

    The wrapper accepts two extra keywords: ``description`` (part of the file name) and ``autoload``
    (``False`` forces a recomputation).
    """
    def decorator_function(function):
        # positional arguments passed by keyword count as well
        posargs_keys = list(inspect.signature(function).parameters)[:nargs]

        @functools.wraps(function)
        def autosave(*args, **kwargs):
            description = kwargs.pop('description', '')
            autoload = kwargs.pop('autoload', True) and load_autosave_data
            if autosave_directory is None:
                return function(*args, **kwargs)
            relevant_args = list(args[:nargs])
            if kwargs_keys is not None:
                for key in [*posargs_keys, *kwargs_keys]:
                    if key in kwargs:
                        relevant_args.append(kwargs[key])
            legacy_csum = checksum(function, *relevant_args)
            csum = legacy_csum if version is None else checksum(version, *relevant_args)
            filename = get_filename(function, csum, description or '', *relevant_args)
            if autoload and (verify_file(filename, csum) or verify_file(filename, legacy_csum)):
                return load_data(filename)
            result = function(*args, **kwargs)
            save_data(filename, csum, result)
            return result
        return autosave
    return decorator_function
