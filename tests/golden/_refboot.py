"""Import the read-only reference (`/root/reference/src/dyban`) inside THIS container only.

Used by `make_golden.py` to mint the golden vectors under `tests/golden/`.  Nothing in `tests/`,
`bench.py` or `__graft_entry__.py` imports this module at run time: `/root/reference` does not exist on
the GPU box.  Recipe = SURVEY.md Appendix E (matplotlib stub + np.asscalar shim, import from a writable
CWD because `dyban/scores.py:13` opens ./output.log at import time).
"""
import os
import sys
import types
import tempfile

import numpy as np

REF_SRC = '/root/reference/src'
REF_DATA = '/root/reference/examples/data'


def import_reference():
    os.environ.setdefault('TQDM_DISABLE', '1')
    sys.dont_write_bytecode = True
    for m in ('matplotlib', 'matplotlib.pyplot'):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if not hasattr(np, 'asscalar'):
        np.asscalar = lambda a: np.asarray(a).item()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix='dyban_ref_'))
    try:
        import dyban  # noqa: F401
        import dyban.network  # noqa: F401
    finally:
        os.chdir(cwd)
    return sys.modules['dyban']


def read_yeast():
    """examples/yeast_tests.py:35-47 — header row is mis-read as NaN and deleted; list = [on, off]."""
    on = np.delete(np.genfromtxt(os.path.join(REF_DATA, 'datayeaston.txt'), delimiter=','), 0, 0)
    off = np.delete(np.genfromtxt(os.path.join(REF_DATA, 'datayeastoff.txt'), delimiter=','), 0, 0)
    return [on, off]


def read_mtor():
    """examples/mTOR_tests.py:36-40 — pandas.read_csv(header row) -> 120 x 11."""
    return [np.genfromtxt(os.path.join(REF_DATA, 'gp_mTOR.csv'), delimiter=',', skip_header=1)]
