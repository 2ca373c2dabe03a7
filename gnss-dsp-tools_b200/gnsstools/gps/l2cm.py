"""GPS L2 CM code (IS-GPS-200): 27-stage Galois register, polynomial 0o445112474, per-PRN
initial state, 10230 chips. Surface of reference gnsstools/gps/l2cm.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 511500
code_length = 10230
_POLY = 0o445112474

l2cm_init = _g.icd_table('gps.l2cm', 'l2cm_init')
l2cm_end_state = _g.icd_table('gps.l2cm', 'l2cm_end_state')

codes = {}


def make_l2cm(prn):
    return _g.lfsr_galois_lsb(_POLY, l2cm_init[prn], code_length)[0]


def l2cm_code(prn):
    if prn not in codes:
        codes[prn] = make_l2cm(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(l2cm_code(prn), chips, frac, incr, n)

rz = np.array([1.0, 0.0])


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator with the L2C return-to-zero time multiplex (out of the acquisition path)."""
    return _g.correlate_sub2(x, chips, frac, incr, c, code_length, rz)


def test_end_state(prn):
    """Register state after code_length-1 steps (IS-GPS-200H end-state column)."""
    return _g.lfsr_galois_lsb(_POLY, l2cm_init[prn], code_length - 1)[1]


if __name__ == '__main__':
    for prn in l2cm_end_state:
        if test_end_state(prn) != l2cm_end_state[prn]:
            print('prn %d: ***mismatch***' % prn)
