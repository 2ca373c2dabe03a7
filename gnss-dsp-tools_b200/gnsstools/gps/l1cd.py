"""GPS L1C data (L1Cd) ranging code (IS-GPS-800): Weil code from the length-10223 Legendre sequence
with the 7-chip expansion inserted at the PRN's insertion point; 10230 chips.
Surface of reference gnsstools/gps/l1cd.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 10230

l1cd_params = _g.icd_table('gps.l1cd', 'l1cd_params')      # prn -> (weil index w, insertion point p)
N = 10223
L = _g.legendre_sequence(N)
_EXPANSION = np.array([0, 1, 1, 0, 1, 0, 0])

codes = {}


def l1cd(prn):
    w, p = l1cd_params[prn]
    W = _g.weil(L, w)
    return np.concatenate((W[:p - 1], _EXPANSION, W[p - 1:]))


def l1cd_code(prn):
    if prn not in codes:
        codes[prn] = l1cd(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(l1cd_code(prn), chips, frac, incr, n)

boc11 = np.array([1.0, -1.0])


def correlate(x, prn, chips, frac, incr, c, boc11):
    """Tracking correlator with BOC(1,1) (out of the acquisition path)."""
    return _g.correlate_sub2(x, chips, frac, incr, c, code_length, boc11)
