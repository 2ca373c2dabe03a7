"""Batched tracking correlators (gnssacq_correlate_epl) against the reference's own
`<sig>.correlate` loops: golden vectors made from /root/reference (tests/golden/make_correlate_golden.py)
— plain C/A, BOC(1,1) L1Cd, RZ-slotted L2CM, CBOC E1b, TMBOC L1Cp. A single wrong chip index would move
a sum by 2|x_i| ~ 1e-4 of its magnitude; the tolerance is 1e-10, so the indices are the reference's."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from make_correlate_golden import CASES      # noqa: E402  (names and modes only)

G = np.load(os.path.join(HERE, 'golden', 'correlate_epl.npz'))


def params_of(name, mode):
    if mode == 0:
        return None
    sub = G[name + '_sub']
    p = [sub[0], sub[1], 0.953463, 0.301511]                  # CBOC weights of gnsstools/galileo/e1b.py:52
    if mode == 3:
        p += list(G[name + '_pattern'])
    return np.array(p, np.float64)


def check(eng):
    for name, _, _, _, mode, n, starts, incr in CASES:
        got = eng.correlate_epl(G[name + '_x'], G[name + '_chips'], G[name + '_start'], float(G[name + '_incr']), mode=mode,
                                params=params_of(name, mode))
        want = G[name + '_want']
        scale = np.sum(np.abs(G[name + '_x']))
        assert np.max(np.abs(got - want)) <= 1e-10 * scale, (name, got, want)
    # several blocks and codes in one call: hypothesis h = (block xsel[h], code csel[h])
    x = np.stack([G['plain_ca_x'], G['plain_ca_x'][::-1].copy()])
    c = np.stack([G['plain_ca_chips'], 1 - G['plain_ca_chips']])
    st = G['plain_ca_start'][:3]
    got = eng.correlate_epl(x, c, np.concatenate([st, st]), float(G['plain_ca_incr']), xsel=[0, 0, 0, 1, 1, 1], csel=[0, 1, 0, 0, 0, 1])
    want = G['plain_ca_want']
    assert abs(got[0] - want[0]) <= 1e-10 * np.sum(np.abs(x[0])) and abs(got[1] + want[1]) <= 1e-10 * np.sum(np.abs(x[0]))
    with pytest.raises(ValueError):
        eng.correlate_epl(x, c, st, 0.25, xsel=[0, 0, 2])     # block index out of range


def test_correlate_epl_on_emulated_kernels():
    import emu_util
    eng = emu_util.emu_engine()
    check(eng)
    eng.close()


def test_host_correlate_wrappers_match_reference_vectors():
    """The module-level correlate() kept for the tracking scripts (host, Numba) against the same vectors."""
    import gnsstools.gps.ca as ca
    import gnsstools.galileo.e1b as e1b
    import gnsstools.gps.l1cp as l1cp
    for name, mod, acc, prn, extra in (('plain_ca', ca, 'ca_code', 7, ()), ('cboc_e1b', e1b, 'e1b_code', 11, (e1b.boc11,)),
                                       ('tmboc_l1cp', l1cp, 'l1cp_code', 4, (l1cp.boc11,))):
        chips = getattr(mod, acc)(prn)
        for st, want in zip(G[name + '_start'], G[name + '_want']):
            got = mod.correlate(G[name + '_x'], prn, 0, float(st), float(G[name + '_incr']), chips, *extra)
            assert abs(got - want) <= 1e-9 * abs(want) + 1e-6, (name, got, want)


@pytest.mark.gpu
def test_correlate_epl_on_gpu():
    from gnsstools import _native
    eng = _native.Engine(0)
    check(eng)
    eng.close()
