"""Galileo E1-B memory code (4092 chips), tabulated per PRN in the ICD and carried bit-packed in
_data/memory_codes.npz. Surface of reference gnsstools/galileo/e1b.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 4092

_table = None
codes = {}


def e1b_code(prn):
    """0/1 chips; KeyError for a PRN the ICD does not define."""
    global _table
    if prn not in codes:
        if _table is None:
            _table = _g.memory_codes('galileo.e1b')
        codes[prn] = _table[prn]
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(e1b_code(prn), chips, frac, incr, n)

boc11 = np.array([1.0, -1.0])


def correlate(x, prn, chips, frac, incr, c, boc11):
    """Tracking correlator with the E1 CBOC subcarrier (out of the acquisition path)."""
    return _g.correlate_cboc(x, chips, frac, incr, c, code_length, boc11, 0.953463, 0.301511)
