"""BeiDou B2b-Q memory code (10230 chips), tabulated per PRN in the ICD and carried bit-packed in
_data/memory_codes.npz. Surface of reference gnsstools/beidou/b2bq.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 10230000
code_length = 10230

_table = None
codes = {}


def b2bq_code(prn):
    """0/1 chips; KeyError for a PRN the ICD does not define."""
    global _table
    if prn not in codes:
        if _table is None:
            _table = _g.memory_codes('beidou.b2bq')
        codes[prn] = _table[prn]
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(b2bq_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)


def accum(x, cp, incr, a, code_length):
    """Chip-bin accumulation a[int(cp)] += x[i] (reference beidou/b2bq.py:65-71; analysis helper)."""
    for i in range(len(x)):
        a[int(cp)] += x[i]
        cp = (cp + incr) % code_length
