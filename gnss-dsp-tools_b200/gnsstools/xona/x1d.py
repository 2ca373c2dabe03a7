"""Xona Pulsar X1D memory code (1023 chips), tabulated per PRN in the ICD and carried bit-packed in
_data/memory_codes.npz. Surface of reference gnsstools/xona/x1d.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 1023

_table = None
codes = {}


def x1d_code(prn):
    """0/1 chips; KeyError for a PRN the ICD does not define."""
    global _table
    if prn not in codes:
        if _table is None:
            _table = _g.memory_codes('xona.x1d')
        codes[prn] = _table[prn]
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(x1d_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
