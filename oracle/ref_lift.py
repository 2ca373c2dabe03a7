"""TEST INFRASTRUCTURE ONLY — run the reference's own search() in this container.

The reference acquire-*.py scripts have no import guard and define search()
locally, so it is lifted verbatim at run time (regex on `def search` ... `return`)
and exec'd with the reference's own gnsstools modules injected. Needs
/root/reference (present only in the build container); used by
tests/golden/make_golden.py and tests/test_oracle_vs_reference.py.
"""

import importlib
import os
import re
import sys

REF = os.environ.get('GNSS_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF, 'acquire-gps-l1.py'))


def ref_import(modname):
    """Import a module of the reference's gnsstools package (e.g. 'gnsstools.gps.ca')
    without letting it shadow the repo's own gnsstools."""
    saved = {k: v for k, v in sys.modules.items() if k == 'gnsstools' or k.startswith('gnsstools.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        cache = _REF_MODULES
        for k, v in cache.items():
            sys.modules[k] = v
        mod = importlib.import_module(modname)
        for k, v in list(sys.modules.items()):
            if k == 'gnsstools' or k.startswith('gnsstools.'):
                cache[k] = v
        return mod
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'gnsstools' or k.startswith('gnsstools.')]:
            del sys.modules[k]
        sys.modules.update(saved)


_REF_MODULES = {}


def lift_search(script):
    """Return the reference search() of acquire-<script>.py as a callable."""
    import numpy as np
    import scipy.fftpack as fft
    src = open(os.path.join(REF, 'acquire-%s.py' % script)).read()
    m = re.search(r'^def search\(.*?^  return [^\n]*\n', src, re.S | re.M)
    if m is None:
        raise RuntimeError('no search() in acquire-%s.py' % script)
    ns = {'np': np, 'fft': fft, 'nco': ref_import('gnsstools.nco')}
    for imp in re.finditer(r'^import gnsstools\.(\w+)\.(\w+) as (\w+)', src, re.M):
        ns[imp.group(3)] = ref_import('gnsstools.%s.%s' % (imp.group(1), imp.group(2)))
    exec(compile(m.group(0), 'acquire-%s.py:search' % script, 'exec'), ns)
    return ns['search'], ns
