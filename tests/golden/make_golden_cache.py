"""Golden fixtures for the on-disk caches, written by the REAL reference (build container only).

    python tests/golden/make_golden_cache.py

* ``checksum.json``       -- ``mdevaluate.checksum.checksum`` of source-independent arguments
                             (160-bit integers, stored as decimal strings);
* ``ref.nojump.npz``      -- the jump-matrix cache ``reader.save_nojump_matrixes`` writes for the
                             ``walk_small`` trajectory (reader.py:205-213), byte for byte;
* ``ref_autosave.npz``    -- the result file ``autosave.autosave_data`` writes for a versioned function
                             (autosave.py:117-127, 142-183) and ``autosave.json`` with its file name.

The reference is imported with the stub modules of make_golden.py; nothing here runs on the GPU box.
"""
import json
import os
import shutil
import sys
import tempfile
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


class Named:
    """An object that brings its own checksum (the protocol Coordinates / readers use)."""

    def __init__(self, value):
        self.value = value

    def __checksum__(self):
        return self.value

    def __call__(self, *args, **kwargs):
        return None


def checksum_cases():
    """name -> argument tuple; shared with tests/test_caches.py (imported from there)."""
    arr = np.arange(12, dtype=np.int32).reshape(3, 4)
    return {
        'int': (1,),
        'str': ('42',),
        'int_as_str': (42,),
        'bytes': (b'\x00\x01abc',),
        'float': (0.25,),
        'none': (None,),
        'many': ('shifted_correlation-1', 10, 0.5, None, False),
        'array_i32': (arr,),
        'array_f64': (np.linspace(0, 1, 7),),
        'module': (np,),
        'object': (Named(123456789),),
        'nested': (2016, Named(2 ** 150 + 12345)),
        'slice': (slice(None), slice(24, None)),
        'partial': (partial(Named('isf-1'), q=22.7, box=None),),
        'partial_args': (partial(Named('f'), 3, np.arange(3.0)),),
    }


def main():
    md = mg.load_reference()
    from mdevaluate import autosave as ref_autosave
    from mdevaluate import checksum as ref_checksum
    from mdevaluate import reader as ref_reader

    # ---- checksums ---------------------------------------------------------------------------------
    out = {name: str(ref_checksum.checksum(*args)) for name, args in checksum_cases().items()}

    @ref_checksum.version(3)
    def versioned(x):
        return x
    out['version_decorator'] = str(ref_checksum.checksum(versioned))
    out['strip_comments'] = ref_checksum.strip_comments(STRIP_SOURCE)
    with open(os.path.join(HERE, 'checksum.json'), 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)

    # ---- nojump cache of the walk_small trajectory ----------------------------------------------------
    g = np.load(os.path.join(HERE, 'walk_small.npz'))
    tmp = tempfile.mkdtemp()
    try:
        class FileReader(mg.FakeReader):
            filename = os.path.join(tmp, 'walk_small.xtc')

            def __checksum__(self):
                return 987654321

        from scipy import sparse
        reader = FileReader(g['pos'], g['box'], g['times'])
        T, M = g['pos'].shape[:2]
        coords = md.coordinates.Coordinates(reader, atom_subset=mg.FakeSubset(M))
        jd = ref_reader.parse_jumps(coords)
        reader.nojump_matrixes = tuple(
            sparse.csr_matrix((np.array(m.data), (np.array(m.row), np.array(m.col))), shape=(T, M)) for m in jd
        )
        ref_reader.save_nojump_matrixes(reader)
        written = ref_reader.nojump_filename(reader)
        assert os.path.basename(written) == '.walk_small.xtc.nojump.npz'
        shutil.copy(written, os.path.join(HERE, 'ref.nojump.npz'))

        # ---- autosave file of a versioned function --------------------------------------------------
        ref_autosave.enable('autosave_dir', verbose=False)

        class Frames:
            frames = reader
            description = 'OW'

            def __checksum__(self):
                return 555

        @ref_autosave.autosave_data(nargs=2, kwargs_keys=('scale',), version='toy-1')
        def toy(function, frames, scale=1.0):
            return np.arange(4.0) * scale, np.arange(8.0).reshape(4, 2) * scale

        toy(Named('fn'), Frames(), scale=2.0, description='run1')
        files = os.listdir(os.path.join(tmp, 'autosave_dir'))
        assert files == ['toy_run1_OW.npz'], files
        shutil.copy(os.path.join(tmp, 'autosave_dir', files[0]), os.path.join(HERE, 'ref_autosave.npz'))
        with open(os.path.join(HERE, 'autosave.json'), 'w') as fh:
            json.dump({'filename': files[0],
                       'checksum': str(ref_checksum.checksum('toy-1', Named('fn'), Frames(), 2.0))}, fh, indent=1)
        ref_autosave.disable()
    finally:
        shutil.rmtree(tmp)


STRIP_SOURCE = '''def f(x):
    """Docstring
    over lines."""
    # a comment
    y = x + 1  # trailing stays
    \'\'\'single quoted block\'\'\'
    "string statement"
    return y
'''


if __name__ == '__main__':
    main()
