"""GLONASS high-accuracy (P) code: 25-stage register 1 + x^3 + x^25, output of stage 10, truncated to 5110000 chips. FDMA — one code for every satellite, so code() has no PRN
argument. Surface of reference gnsstools/glonass/p.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 5110000
code_length = 5110000

c = _g.lfsr_fibonacci(25, (24, 2), (1 << 25) - 1, code_length, out_tap=9)


def p_code():
    return c


def code(chips, frac, incr, n):
    return _g.resample(c, chips, frac, incr, n)


def correlate(x, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path)."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)
