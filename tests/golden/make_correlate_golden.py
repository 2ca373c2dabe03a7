"""Golden vectors for the batched tracking correlators (gnssacq_correlate_epl), produced by the
REFERENCE's own `<sig>.correlate` loops (Numba) imported from /root/reference:
gps.ca (plain), gps.l1cd (x BOC(1,1)), gps.l2cm (x RZ slots), galileo.e1b (CBOC), gps.l1cp (TMBOC).
Build container only:  python tests/golden/make_correlate_golden.py -> tests/golden/correlate_epl.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]

from oracle import ref_lift          # noqa: E402

# (name, reference module, code accessor, prn, mode, samples, start phases (chips+frac), incr)
CASES = [
    ('plain_ca', 'gnsstools.gps.ca', 'ca_code', 7, 0, 4097, [100.25 - 0.05, 100.25, 100.25 + 0.05, -0.3, 1022.99], 1023000.0 / 4.0e6 * (1 + 3e-6)),
    ('boc_l1cd', 'gnsstools.gps.l1cd', 'l1cd_code', 3, 1, 5000, [5000.5 - 0.2, 5000.5, 5000.5 + 0.2], 1023000.0 / 5.0e6),
    ('rz_l2cm', 'gnsstools.gps.l2cm', 'l2cm_code', 5, 1, 6001, [10229.75, 17.125, -0.6], 511500.0 / 2.5e6),
    ('cboc_e1b', 'gnsstools.galileo.e1b', 'e1b_code', 11, 2, 8184, [2000.3 - 0.05, 2000.3, 2000.3 + 0.05], 1023000.0 / 8.184e6),
    ('tmboc_l1cp', 'gnsstools.gps.l1cp', 'l1cp_code', 4, 3, 9000, [7000.7 - 0.05, 7000.7, 7000.7 + 0.05, 10229.999], 1023000.0 / 9.0e6 * (1 - 2e-6)),
]


def main():
    out = {}
    rng = np.random.default_rng(99)
    for name, modname, acc, prn, mode, n, starts, incr in CASES:
        mod = ref_lift.ref_import(modname)
        chips = np.array([int(v) for v in getattr(mod, acc)(prn)])      # (l1cd's table holds sympy integers; same values)
        x = (rng.integers(-60, 61, n) + 1j * rng.integers(-60, 61, n)).astype(np.complex64)
        res = []
        for st in starts:
            if mode == 0:
                p = mod.correlate(x, prn, 0, st, incr, chips)
            elif name == 'rz_l2cm':
                p = mod.correlate(x, prn, 0, st, incr, chips)
            else:
                p = mod.correlate(x, prn, 0, st, incr, chips, mod.boc11)
            res.append(complex(p))
        out[name + '_x'] = x
        out[name + '_chips'] = (np.asarray(chips) != 0).astype(np.int8)
        out[name + '_start'] = np.array(starts, np.float64)
        out[name + '_incr'] = np.float64(incr)
        out[name + '_want'] = np.array(res, np.complex128)
        if name == 'rz_l2cm':
            out[name + '_sub'] = np.asarray(mod.rz, np.float64)
        elif mode >= 1:
            out[name + '_sub'] = np.asarray(mod.boc11, np.float64)
        if mode == 3:
            out[name + '_pattern'] = np.asarray(mod.tmboc_pattern, np.float64)
        print(name, res[:2])
    np.savez_compressed(os.path.join(HERE, 'correlate_epl.npz'), **out)


if __name__ == '__main__':
    main()
