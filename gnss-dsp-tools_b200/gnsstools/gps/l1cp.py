"""GPS L1C pilot (L1Cp) ranging code (IS-GPS-800): Weil code from the length-10223 Legendre sequence
with the 7-chip expansion inserted at the PRN's insertion point; 10230 chips.
Surface of reference gnsstools/gps/l1cp.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 10230

l1cp_params = _g.icd_table('gps.l1cp', 'l1cp_params')      # prn -> (weil index w, insertion point p)
N = 10223
L = _g.legendre_sequence(N)
_EXPANSION = np.array([0, 1, 1, 0, 1, 0, 0])

codes = {}


def l1cp(prn):
    w, p = l1cp_params[prn]
    W = _g.weil(L, w)
    return np.concatenate((W[:p - 1], _EXPANSION, W[p - 1:]))


def l1cp_code(prn):
    if prn not in codes:
        codes[prn] = l1cp(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(l1cp_code(prn), chips, frac, incr, n)


# 1800-chip overlay codes: 11-stage register s <- [parity(s & poly)] + s[0:10], output s[10];
# PRN >= 64 xor a second register with polynomial 0o5001 (IS-GPS-800 table 3.2-3).
l1cp_secondary_params = _g.icd_table('gps.l1cp', 'l1cp_secondary_params')
sec_code_length = 1800
secondary_codes = {}


def _overlay(poly, init):
    full, mask = (1 << 11) - 1, poly // 2
    out = np.empty(sec_code_length)
    x = init
    for i in range(sec_code_length):
        out[i] = (x >> 10) & 1
        x = ((x << 1) & full) | (bin(x & mask).count('1') & 1)
    return out


def secondary_code(prn):
    if prn not in secondary_codes:
        par = l1cp_secondary_params[prn]
        c = _overlay(par[0], par[1])
        if prn >= 64:
            c = np.logical_xor(c, _overlay(0o5001, par[2])).astype(np.float64)
        secondary_codes[prn] = c
    return secondary_codes[prn]


boc11 = np.array([1.0, -1.0])
tmboc_pattern = np.array([1, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0])


def correlate(x, prn, chips, frac, incr, c, boc11):
    """Tracking correlator with the TMBOC(6,1,4/33) pattern (out of the acquisition path)."""
    return _g.correlate_tmboc(x, chips, frac, incr, c, code_length, boc11, tmboc_pattern)
