"""Xona Pulsar X1P memory code (1023 chips), tabulated per PRN in the ICD and carried bit-packed in
_data/memory_codes.npz. Surface of reference gnsstools/xona/x1p.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 1023

secondary_code = 1.0 - 2.0 * np.array([int(b) for b in _g.icd_table('xona.x1p', 'secondary_bits')[0]], dtype=np.float64)

_table = None
codes = {}


def x1p_code(prn):
    """0/1 chips; KeyError for a PRN the ICD does not define."""
    global _table
    if prn not in codes:
        if _table is None:
            _table = _g.memory_codes('xona.x1p')
        codes[prn] = _table[prn]
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(x1p_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
