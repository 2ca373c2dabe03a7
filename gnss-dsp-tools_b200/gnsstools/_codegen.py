"""Shared helpers for the PRN code generator modules.

Every generator module exposes the reference surface
``chip_rate, code_length, codes, <sig>_code(prn), code(prn, chips, frac, incr, n)``
(e.g. reference gnsstools/gps/ca.py:101-112); the resampling rule is common and
lives here.
"""

import numpy as np


def resample(chip_bits, chips, frac, incr, n):
    """+-1.0 float64[n]: 1 - 2*c[floor((chips % L) + frac + incr*i) mod L]
    (reference gnsstools/gps/ca.py:106-112; identical body in every module)."""
    L = len(chip_bits)
    ph = (chips % L) + frac + incr * np.arange(n)
    k = np.mod(np.floor(ph).astype('int'), L)
    return 1.0 - 2.0 * chip_bits[k]


def lfsr_fibonacci(nbits, taps, state, length, out_tap=None):
    """Run a Fibonacci LFSR for `length` steps.

    `state` is a list of bits x[0..nbits-1]; each step outputs x[out_tap]
    (default: the last stage), then shifts right inserting the XOR of the
    `taps` stages (0-based) at x[0]. Returns a float64 0/1 array.
    """
    if out_tap is None:
        out_tap = nbits - 1
    x = list(state)
    out = np.empty(length)
    for i in range(length):
        out[i] = x[out_tap]
        fb = 0
        for t in taps:
            fb ^= x[t]
        x = [fb] + x[:-1]
    return out
