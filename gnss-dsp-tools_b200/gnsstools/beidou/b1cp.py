"""BeiDou B1C pilot (B1Cp) ranging code (BDS-SIS-ICD-B1C): Weil code of the length-10243 Legendre
sequence, truncated to 10230 chips from the PRN's truncation point.
Surface of reference gnsstools/beidou/b1cp.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 10230

b1cp_params = _g.icd_table('beidou.b1cp', 'b1cp_params')     # prn -> (phase difference w, truncation point p)
N = 10243
L = _g.legendre_sequence(N)

codes = {}


def b1cp(prn):
    w, p = b1cp_params[prn]
    return _g.weil_truncated(L, w, p, code_length)


def b1cp_code(prn):
    if prn not in codes:
        codes[prn] = b1cp(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(b1cp_code(prn), chips, frac, incr, n)


b1cp_secondary_params = _g.icd_table('beidou.b1cp', 'b1cp_secondary_params')
sec_N = 3607
sec_L = _g.legendre_sequence(sec_N)
sec_code_length = 1800
secondary_codes = {}


def secondary_code(prn):
    """0/1 overlay code: truncated Weil code of length-3607 Legendre sequence."""
    if prn not in secondary_codes:
        w, p = b1cp_secondary_params[prn]
        secondary_codes[prn] = _g.weil_truncated(sec_L, w, p, sec_code_length)
    return secondary_codes[prn]

boc11 = np.array([1.0, -1.0])


def correlate(x, prn, chips, frac, incr, c, boc11):
    """Tracking correlator with BOC(1,1) (out of the acquisition path)."""
    return _g.correlate_sub2(x, chips, frac, incr, c, code_length, boc11)
