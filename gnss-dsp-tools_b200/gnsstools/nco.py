"""Table NCO, BOC(1,1) subcarrier and carrier wipe-off (reference: gnsstools/nco.py).

``nco`` and ``boc11`` are replica/setup helpers and stay numpy on the host.
``mix`` (whole-capture carrier wipe-off, reference nco.py:30-41) runs on the GPU
through ``gnssacq_mix``; it raises if the CUDA library is unavailable.
"""

import numpy as np

NT = 1024
nco_table = np.exp(2j * np.pi * np.arange(NT) * (1.0 / NT))


def nco(f, p, n):
    """complex128[n]: table[floor((p + f*i)*1024) mod 1024] (reference nco.py:6-10)."""
    ph = p + f * np.arange(n)
    k = np.floor(ph * NT).astype('int')
    return nco_table[np.mod(k, NT)]


def boc11(chips, frac, incr, n):
    """+-1 BOC(1,1) square subcarrier sampled at `incr` chips/sample (reference nco.py:12-19)."""
    ph = (chips % 2) + frac + incr * np.arange(n)
    half = np.floor(ph * 2).astype('int')
    return 2 * np.mod(half, 2) - 1


def mix(x, f, p):
    """In-place x[i] *= table[((dp0 + i*df) >> 50) & 1023] with the reference's
    2^50-scaled int64 phase accumulator (reference nco.py:30-41). GPU only."""
    from . import _native
    _native.mix_inplace(x, f, p)
