"""BeiDou B2a pilot (B2ap) ranging code (BDS-SIS-ICD-B2a): two 13-stage registers, G1 from all ones and
re-initialised after chip 8189 (period 8190), G2 from the PRN's initial state; 10230 chips.
Surface of reference gnsstools/beidou/b2ap.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

b2ap_g2_initial = _g.icd_table('beidou.b2ap', 'b2ap_g2_initial')     # prn -> 13-character bit string, stage 1 first

_G1_TAPS = (2, 5, 6, 12)
_G2_TAPS = (0, 4, 6, 7, 11, 12)
_g1 = _g.stage(_g.lfsr_states(13, _G1_TAPS, 0x1fff, code_length, reset_after=8189, reset_to=0x1fff), 12)

codes = {}


def b2ap(prn):
    g2 = _g.lfsr_fibonacci(13, _G2_TAPS, b2ap_g2_initial[prn], code_length)
    return np.logical_xor(_g1, g2).astype(np.float64)


def b2ap_code(prn):
    if prn not in codes:
        codes[prn] = b2ap(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(b2ap_code(prn), chips, frac, incr, n)


b2ap_secondary_params = _g.icd_table('beidou.b2ap', 'b2ap_secondary_params')
sec_N = 1021
sec_L = _g.legendre_sequence(sec_N)
sec_code_length = 100
secondary_codes = {}


def secondary_code(prn):
    """0/1 overlay code: truncated Weil code of length-1021 Legendre sequence."""
    if prn not in secondary_codes:
        w, p = b2ap_secondary_params[prn]
        secondary_codes[prn] = _g.weil_truncated(sec_L, w, p, sec_code_length)
    return secondary_codes[prn]


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
