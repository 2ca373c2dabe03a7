#!/usr/bin/env python
"""Throughput of the correlator bank on the serial long-code searches at the reference's own sizes
(acquire-glonass-l1-p.py: 69.984 Msps, --time 80 -> 1000 hypotheses x 20 blocks x 279936 samples;
acquire-gps-l2cl.py: --time 40 -> 75 x 2 x 1399680), next to the oracle on a bounded sample."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]
from gnsstools import _native, acquire_serial
from oracle import acq_oracle as orc
import gnsstools.glonass.p as gp
import gnsstools.gps.l2cl as l2cl

eng = _native.Engine(0)
rng = np.random.default_rng(0)
fs = 69.984e6
for name, ms in (('glonass-l1-p', 80), ('gps-l2cl', 40)):
    nx = int(fs * 0.001 * (ms + 5))
    x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
    if name == 'glonass-l1-p':
        run = lambda: acquire_serial.search_glonass_p(None, -2, 310.0, 278.6, ms, fs, 562500, engine=eng)
        H, B, n = 1000, ms // 4, int(fs * 0.004)
        cpu = lambda h: orc.search_glonass_p(x, gp.p_code(), fs, 562500, -2, 310.0, 278.6, ms, hypotheses=h)
    else:
        run = lambda: acquire_serial.search_l2cl(None, 3, 431.0, 8317.2, ms, fs, engine=eng)
        H, B, n = 75, ms // 20, int(fs * 0.020)
        cpu = lambda h: orc.search_l2cl(x, l2cl.l2cl_code(3), fs, 431.0, 8317.2, ms, hypotheses=h)
    eng.set_signal(x[:B * n])
    run()
    t0 = time.perf_counter()
    for _ in range(3):
        got = run()
    dt = (time.perf_counter() - t0) / 3
    hs = 4
    t0 = time.perf_counter()
    want = cpu(hs)
    dc = (time.perf_counter() - t0) * H / hs
    print('%-14s H=%d B=%d n=%d: GPU %.2f ms per search (host call to host result) = %.3e sample-hypotheses/s; '
          'oracle (1 core, %d hypotheses extrapolated) %.1f s' % (name, H, B, n, dt * 1e3, H * B * n / dt, hs, dc), flush=True)
