"""Galileo E6-B memory code (5115 chips), tabulated per PRN in the ICD and carried bit-packed in
_data/memory_codes.npz. Surface of reference gnsstools/galileo/e6b.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 5115000
code_length = 5115

_table = None
codes = {}


def e6b_code(prn):
    """0/1 chips; KeyError for a PRN the ICD does not define."""
    global _table
    if prn not in codes:
        if _table is None:
            _table = _g.memory_codes('galileo.e6b')
        codes[prn] = _table[prn]
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(e6b_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
