"""CPU checks of the boundary: the library loads and exports every symbol include/dbn_b200.h declares."""
import os
import re

import pytest

from helpers import ROOT


def test_header_symbols_exported():
    from dynamicbayesiannetworks_b200 import build, _lib
    build.build()
    header = open(os.path.join(ROOT, 'include', 'dbn_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|void|const char\*)\s+(dbn_[a-z0-9_]+)\s*\(', header, flags=re.M))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()          # resolves every symbol; raises if one is missing
    assert lib.dbn_abi_version() == 1


def test_layout_constants_match_header():
    from dynamicbayesiannetworks_b200 import _lib as L
    header = open(os.path.join(ROOT, 'include', 'dbn_b200.h')).read()
    env = {}
    for name, expr in re.findall(r'^#define\s+(DBN_[A-Z0-9_]+)\s+(.+?)\s*(?:/\*.*)?$', header, flags=re.M):
        try:
            env[name] = int(eval(expr.replace('/', '//'), {}, env))
        except Exception:
            pass
    assert env['DBN_RD_STRIDE'] == L.RD_STRIDE and env['DBN_RI_STRIDE'] == L.RI_STRIDE
    assert env['DBN_TD_STRIDE'] == L.TD_STRIDE and env['DBN_TI_STRIDE'] == L.TI_STRIDE
    for k in ('RD_MU_PI', 'RD_MU_TAU', 'RD_BETA', 'TD_MUS_COV', 'TD_BETA_MEAN', 'TD_BETA_COV', 'TD_BETA', 'TD_MU',
              'TD_MUS_LOGDENS', 'TD_MUS_MEAN', 'TI_TAU', 'TI_S', 'TI_PI', 'TI_S_BEFORE'):
        assert env['DBN_' + k] == getattr(L, k), k


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly, not compute on the CPU."""
    from helpers import has_gpu
    if has_gpu():
        pytest.skip('GPU present')
    from dynamicbayesiannetworks_b200.engine import DbnEngine, DbnError
    with pytest.raises(DbnError):
        DbnEngine(n_chains=1)
