"""The C-ABI library loads (no GPU needed) and exports every symbol include/gnssacq.h declares;
on a machine without a CUDA device the product path raises instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'gnssacq.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gnssacq_[a-z_0-9]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from gnsstools import _native
    return ctypes.CDLL(_native.LIB_PATH)


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert 'gnssacq_search' in names and 'gnssacq_mix' in names and len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header():
    from gnsstools import _native
    import inspect
    src = inspect.getsource(_native)
    for n in declared_symbols():
        assert n in src, 'ctypes binding lacks ' + n


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from gnsstools import _native
    with pytest.raises((_native.NativeError, ValueError)):
        _native.Engine(0)
