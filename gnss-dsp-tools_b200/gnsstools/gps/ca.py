"""GPS L1 C/A code (IS-GPS-200): G1 xor delayed G2, 1023 chips, PRN 1-210.
Surface of reference gnsstools/gps/ca.py (chip_rate, code_length, g2_delay, codes, ca_code, code)."""

import numpy as np

from .. import _codegen as _g

chip_rate = 1023000
code_length = 1023

# G2 delay (chips) for PRN 1..210, IS-GPS-200H tables 3-Ia/3-Ib and 6-I.
_G2_DELAY = (
    5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257,
    258, 469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860,
    861, 862, 863, 950, 947, 948, 950, 67, 103, 91, 19, 679, 225, 625, 946,
    638, 161, 1001, 554, 280, 710, 709, 775, 864, 558, 220, 397, 55, 898, 759,
    367, 299, 1018, 729, 695, 780, 801, 788, 732, 34, 320, 327, 389, 407, 525,
    405, 221, 761, 260, 326, 955, 653, 699, 422, 188, 438, 959, 539, 879, 677,
    586, 153, 792, 814, 446, 264, 1015, 278, 536, 819, 156, 957, 159, 712, 885,
    461, 248, 713, 126, 807, 279, 122, 197, 693, 632, 771, 467, 647, 203, 145,
    175, 52, 21, 237, 235, 886, 657, 634, 762, 355, 1012, 176, 603, 130, 359,
    595, 68, 386, 797, 456, 499, 883, 307, 127, 211, 121, 118, 163, 628, 853,
    484, 289, 811, 202, 1021, 463, 568, 904, 670, 230, 911, 684, 309, 644, 932,
    12, 314, 891, 212, 185, 675, 503, 150, 395, 345, 846, 798, 992, 357, 995,
    877, 112, 144, 476, 193, 109, 445, 291, 87, 399, 292, 901, 339, 208, 711,
    189, 263, 537, 663, 942, 173, 900, 30, 500, 935, 556, 373, 85, 652, 310,
)
g2_delay = {prn: d for prn, d in enumerate(_G2_DELAY, start=1)}

# G1: 1 + x^3 + x^10 ; G2: 1 + x^2 + x^3 + x^6 + x^8 + x^9 + x^10 ; all-ones start.
g1 = _g.lfsr_fibonacci(10, (2, 9), 0x3ff, code_length)
g2 = _g.lfsr_fibonacci(10, (1, 2, 5, 7, 8, 9), 0x3ff, code_length)

codes = {}


def ca_code(prn):
    """0/1 chips of C/A PRN `prn`; KeyError for an unknown PRN (as the reference)."""
    if prn not in codes:
        codes[prn] = np.logical_xor(g1, np.roll(g2, g2_delay[prn]))
    return codes[prn]


def code(prn, chips, frac, incr, n):
    return _g.resample(ca_code(prn), chips, frac, incr, n)


def correlate(x, prn, chips, frac, incr, c):
    """Tracking correlator (out of the acquisition path); see _codegen.correlate_plain."""
    return _g.correlate_plain(x, chips, frac, incr, c, code_length)


def first_10_chips(prn):
    """First ten chips as an integer (IS-GPS-200H octal check column)."""
    return int(''.join('%d' % b for b in ca_code(prn)[:10]), 2)


if __name__ == '__main__':
    for prn in g2_delay:
        print('%d: %04o' % (prn, first_10_chips(prn)))
