"""GPS L2 CL code (IS-GPS-200): 27-stage Galois register, polynomial 0o445112474, per-PRN
initial state, 767250 chips. Surface of reference gnsstools/gps/l2cl.py."""

import numpy as np

from .. import _codegen as _g

chip_rate = 511500
code_length = 767250
_POLY = 0o445112474

l2cl_init = _g.icd_table('gps.l2cl', 'l2cl_init')
l2cl_end_state = _g.icd_table('gps.l2cl', 'l2cl_end_state')

codes = {}


def make_l2cl(prn):
    return _g.lfsr_galois_lsb(_POLY, l2cl_init[prn], code_length)[0]


def l2cl_code(prn):
    if prn not in codes:
        codes[prn] = make_l2cl(prn)
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(l2cl_code(prn), chips, frac, incr, n)

rz = np.array([1.0, 0.0])


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator with the L2C return-to-zero time multiplex (out of the acquisition path)."""
    return _g.correlate_sub2(x, chips, frac, incr, c, code_length, rz)


def test_end_state(prn):
    """Register state after code_length-1 steps (IS-GPS-200H end-state column)."""
    return _g.lfsr_galois_lsb(_POLY, l2cl_init[prn], code_length - 1)[1]


if __name__ == '__main__':
    for prn in l2cl_end_state:
        if test_end_state(prn) != l2cl_end_state[prn]:
            print('prn %d: ***mismatch***' % prn)
