"""BeiDou B1C data (B1Cd) ranging code (BDS-SIS-ICD-B1C): Weil code of the length-10243 Legendre
sequence, truncated to 10230 chips from the PRN's truncation point.
Surface of reference gnsstools/beidou/b1cd.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 10230

b1cd_params = _g.icd_table('beidou.b1cd', 'b1cd_params')     # prn -> (phase difference w, truncation point p)
N = 10243
L = _g.legendre_sequence(N)

codes = {}


def b1cd(prn):
    w, p = b1cd_params[prn]
    return _g.weil_truncated(L, w, p, code_length)


def b1cd_code(prn):
    if prn not in codes:
        codes[prn] = b1cd(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(b1cd_code(prn), chips, frac, incr, n)

boc11 = np.array([1.0, -1.0])


def correlate(x, prn, chips, frac, incr, c, boc11):
    """Tracking correlator with BOC(1,1) (out of the acquisition path)."""
    return _g.correlate_sub2(x, chips, frac, incr, c, code_length, boc11)
