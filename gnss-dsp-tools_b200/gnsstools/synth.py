"""Synthetic captures for tests and the benchmark (not part of the reference surface).

capture() returns what a script's search() receives: complex samples at the script's
internal rate, (ms+5) ms long (acquire-gps-l1.py:80-83), Gaussian noise (sigma 8 per rail)
plus planted satellites, rounded to integers and clipped to +-127 like an 8-bit recording.
"""

import numpy as np

from . import acquire as _acq
from ._codegen import resample


def capture(signal, ms, sats=(), seed=0, extra_ms=5, noise=8.0):
    """sats: iterable of (key, doppler_hz, code_phase_chips, amplitude)."""
    sig = _acq.SIGNALS[signal] if isinstance(signal, str) else signal
    mod = _acq.code_module(sig)
    rng = np.random.default_rng(seed)
    per_ms = sig.fs * 0.001
    nx = int(round(per_ms * (ms + extra_ms)))
    x = rng.normal(0.0, noise, nx) + 1j * rng.normal(0.0, noise, nx)
    t = np.arange(nx)
    L = mod.code_length
    incr = float(sig.periods * L) / sig.n    # chips per sample at the script's internal rate
    for key, doppler, phase, amp in sats:
        name = sig.module.split('.')[-1] + '_code'
        chips = getattr(mod, name)() if sig.fdma else getattr(mod, name)(key)
        c = resample(np.asarray(chips, dtype=np.float64), phase, 0, incr, nx)
        if sig.boc:
            from . import nco
            c = c * nco.boc11(phase, 0, incr, nx)
        fc = doppler + (sig.carrier_step * key if sig.fdma else 0.0)
        x += amp * c * np.exp(2j * np.pi * fc * t / sig.fs)
    x = np.clip(np.round(x.real), -127, 127) + 1j * np.clip(np.round(x.imag), -127, 127)
    return x.astype(np.complex64)
