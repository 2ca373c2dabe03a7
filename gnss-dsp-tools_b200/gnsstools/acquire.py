"""Batched FFT acquisition search on the GPU — the B200 replacement for the search() /
worker() / mp.Pool fan-out that every reference acquire-*.py script carries
(acquire-gps-l1.py:18-40,100-108 and the variants in SURVEY.md §8a / Appendix A).

``acquire(signal, x, keys, doppler_search, ms)`` runs every (PRN x Doppler bin x block) of a
script's search in one call into libgnssacq.so; ``search(signal, x, key, doppler_search, ms)``
keeps the reference's per-PRN signature and return value ``(metric, code_chips, doppler_hz)``.
All arithmetic of the search runs on the device, including the replica set-up (resampled code,
BOC(1,1), zero half, FFT) from the chip tables; this module only hands over the tables and
converts the returned lag index to chips.
"""

import importlib

import numpy as np

from . import nco
from . import _native


class Signal:
    """Constants hard-coded inside one reference script's search()."""

    def __init__(self, module, fs, n, blocks, pad=False, boc=False, normalize=False, mod_L=False,
                 carrier_step=0.0, fdma=False, fmt=None, prns='1-32', doppler='-7000,7000,200',
                 cutoff=None, time=80, periods=1):
        self.module, self.fs, self.n, self.blocks = module, fs, n, blocks
        self.pad, self.boc, self.normalize, self.mod_L = pad, boc, normalize, mod_L
        self.carrier_step, self.fdma = carrier_step, fdma
        self.fmt, self.prns, self.doppler, self.cutoff, self.time = fmt, prns, doppler, cutoff, time
        self.periods = periods        # code periods per coherent block (1 in every reference script)

    @property
    def N(self):
        return 2 * self.n if self.pad else self.n


_ms = lambda ms: ms
_F1 = 'prn %3d doppler % 7.1f metric % 5.2f code_offset %6.1f'
_F2 = 'prn %2d doppler % 7.1f metric % 7.1f code_offset %6.1f'
_F3 = 'prn %3d doppler % 7.1f metric % 7.1f code_offset %6.1f'
_FG = 'chan % 2d doppler % 7.1f metric % 7.1f code_offset %7.2f'

SIGNALS = {
    # variant A: circular, metric normalised by mean(q)   (acquire-gps-l1.py:18-40)
    'gps-l1': Signal('gps.ca', 4096000.0, 4096, _ms, normalize=True, fmt=_F1, cutoff=1.5e6),
    'xona-x1': Signal('xona.x1p', 4096000.0, 4096, _ms, normalize=True, fmt=_F1, prns='0',
                      doppler='-50000,50000,200', cutoff=1.5e6),
    'xona-x5p': Signal('xona.x5p', 30690000.0, 30690, _ms, normalize=True, fmt=_F1, prns='0',
                       doppler='-50000,50000,200', cutoff=12e6),
    # variant B: circular, raw metric, FDMA channels       (acquire-glonass-l1.py:18-39)
    'glonass-l1': Signal('glonass.ca', 16384000.0, 16384, _ms, carrier_step=562500.0, fdma=True, fmt=_FG,
                         prns='-7:7', cutoff=6e6),
    'glonass-l2': Signal('glonass.ca', 16384000.0, 16384, _ms, carrier_step=437500.0, fdma=True, fmt=_FG,
                         prns='-7:7', cutoff=6e6),
    # variant C: circular + BOC(1,1), 10 ms coherent       (acquire-gps-l1cd.py:18-42)
    'gps-l1cd': Signal('gps.l1cd', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True, fmt=_F3,
                       doppler='-7000,7000,20', cutoff=4e6),
    'gps-l1cp': Signal('gps.l1cp', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True, fmt=_F3,
                       doppler='-7000,7000,20', cutoff=4e6),
    'beidou-b1cd': Signal('beidou.b1cd', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True, fmt=_F3,
                          prns='1-63', doppler='-7000,7000,20', cutoff=4e6),
    'beidou-b1cp': Signal('beidou.b1cp', 8192000.0, 81920, lambda ms: ms // 10, boc=True, mod_L=True, fmt=_F3,
                          prns='1-63', doppler='-7000,7000,20', cutoff=4e6),
    # variant D: zero-padded 2n, half-overlapping blocks   (acquire-gps-l5i.py:18-40, acquire-galileo-e1b.py:18-42)
    'galileo-e1b': Signal('galileo.e1b', 8192000.0, 32768, lambda ms: ms // 4 - 1, pad=True, boc=True, mod_L=True,
                          fmt=_F2, prns='1-50', doppler='-9000,9000,50', cutoff=4e6),
    'galileo-e1c': Signal('galileo.e1c', 8192000.0, 32768, lambda ms: ms // 4 - 1, pad=True, boc=True, mod_L=True,
                          fmt=_F2, prns='1-50', doppler='-9000,9000,50', cutoff=4e6),
    'beidou-b1i': Signal('beidou.b1i', 8192000.0, 8192, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-63', cutoff=3e6),
    'beidou-b2i': Signal('beidou.b1i', 8192000.0, 8192, _ms, pad=True, mod_L=True, fmt=_F3, prns='1-63', cutoff=3e6),
    'gps-l2cm': Signal('gps.l2cm', 4096000.0, 81920, lambda ms: ms // 20 - 1, pad=True, mod_L=True, fmt=_F3,
                       doppler='-7000,7000,20', cutoff=1.5e6),
    'gps-l5i': Signal('gps.l5i', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, cutoff=12e6),
    'gps-l5q': Signal('gps.l5q', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, cutoff=12e6),
    'galileo-e5ai': Signal('galileo.e5ai', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-50',
                           doppler='-9000,9000,200', cutoff=12e6),
    'galileo-e5aq': Signal('galileo.e5aq', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-50',
                           doppler='-9000,9000,200', cutoff=12e6),
    'galileo-e5bi': Signal('galileo.e5bi', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-50',
                           doppler='-9000,9000,200', cutoff=12e6),
    'galileo-e5bq': Signal('galileo.e5bq', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-50',
                           doppler='-9000,9000,200', cutoff=12e6),
    'galileo-e6b': Signal('galileo.e6b', 15345000.0, 15345, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-50',
                          doppler='-9000,9000,200', cutoff=6e6),
    'galileo-e6c': Signal('galileo.e6c', 15345000.0, 15345, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-50',
                          doppler='-9000,9000,200', cutoff=6e6),
    'beidou-b2ad': Signal('beidou.b2ad', 30690000.0, 30690, lambda ms: 80, pad=True, mod_L=True, fmt=_F3,
                          prns='1-63', cutoff=12e6),
    'beidou-b2ap': Signal('beidou.b2ap', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F3, prns='1-63', cutoff=12e6),
    'beidou-b2bi': Signal('beidou.b2bi', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F3, prns='', cutoff=12e6),
    'beidou-b2bq': Signal('beidou.b2bq', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F3, prns='', cutoff=12e6),
    'beidou-b3i': Signal('beidou.b3i', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, prns='1-63', cutoff=12e6),
    'glonass-l3ocd': Signal('glonass.l3ocd', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, prns='0-63', cutoff=12e6),
    'glonass-l3ocp': Signal('glonass.l3ocp', 30690000.0, 30690, _ms, pad=True, mod_L=True, fmt=_F2, prns='0-63', cutoff=12e6),
}


def code_module(sig):
    return importlib.import_module('gnsstools.' + sig.module)


def replica(sig, key):
    """Time-domain replica the reference hands to fft.fft(): resampled code, optional
    BOC(1,1), optional zero half (acquire-gps-l1.py:22-24, acquire-galileo-e1b.py:23-26)."""
    mod = code_module(sig)
    incr = float(sig.periods * mod.code_length) / sig.n
    c = mod.code(0, 0, incr, sig.n) if sig.fdma else mod.code(key, 0, 0, incr, sig.n)
    if sig.boc:
        c = c * nco.boc11(0, 0, incr, sig.n)
    out = np.zeros(sig.N, dtype=np.int8)          # +-1 and 0 exactly; int8 keeps the upload small
    out[:sig.n] = c
    return out


_chip_cache = {}


def chip_table(sig, key):
    """0/1 chips of one code period, read through the module's public code() at one sample per
    chip (every generator names its table accessor differently; code() is common to all)."""
    ck = (sig.module, None if sig.fdma else key)
    if ck not in _chip_cache:
        mod = code_module(sig)
        L = mod.code_length
        c = mod.code(0, 0, 1.0, L) if sig.fdma else mod.code(key, 0, 0, 1.0, L)
        _chip_cache[ck] = ((1.0 - c) * 0.5).astype(np.int8)
    return _chip_cache[ck]


def set_replicas(eng, sig, keys):
    """Replicas of `keys` built on the device from the chip tables (gnssacq_set_replicas_from_chips):
    resampling, BOC(1,1) and zero padding of replica() without the per-sample host work and upload."""
    mod = code_module(sig)
    incr = float(sig.periods * mod.code_length) / sig.n
    chips = np.stack([chip_table(sig, k) for k in keys])
    eng.set_replicas_from_chips(chips, sig.n, sig.N, incr, boc=sig.boc)


def doppler_bins(doppler_search):
    lo, hi, step = doppler_search
    return np.arange(lo, hi, step)          # max excluded (acquire-gps-l1.py:26)


def _finish(sig, L, bins, metric, lag, dbin):
    if dbin < 0:
        m_metric, m_code, m_doppler = 0, 0, 0          # nothing exceeded 0 (acquire-gps-l1.py:25)
    else:
        m_metric = float(metric)
        m_code = (sig.periods * L) * (float(int(lag)) / sig.n)   # acquire-gps-l1.py:38
        m_doppler = bins[dbin]
    if sig.mod_L:
        m_code = m_code % L                             # acquire-gps-l5i.py:39
    return m_metric, m_code, m_doppler


def acquire(signal, x, keys, doppler_search, ms, engine=None, lag_limit=0, blocks=None):
    """Search every key (PRN, or FDMA channel for GLONASS L1/L2) of `signal` in capture `x`
    (complex, already at the script's internal rate; None = the capture the engine already
    holds on the device). Returns [(metric, code_chips, doppler_hz)]
    in the order of `keys`, each equal to what the reference search() returns."""
    sig = SIGNALS[signal] if isinstance(signal, str) else signal
    eng = engine if engine is not None else _native.default_engine()
    keys = list(keys)
    if not keys:
        return []
    B = sig.blocks(ms) if blocks is None else blocks
    L = code_module(sig).code_length
    bins = doppler_bins(doppler_search)
    if B <= 0 or len(bins) == 0:
        return [_finish(sig, L, bins, 0.0, 0, -1) for _ in keys]
    need = (B - 1) * sig.n + sig.N
    if x is None:
        # capture already resident on the device (acquire_cli.preprocess / Engine.preprocess)
        if getattr(eng, 'n_samples', 0) < need:
            raise ValueError('resident capture too short: %d samples, search needs %d' % (getattr(eng, 'n_samples', 0), need))
    else:
        x = np.asarray(x)
        if x.shape[0] < need:
            raise ValueError('capture too short: %d samples, search needs %d' % (x.shape[0], need))
        eng.set_signal(np.ascontiguousarray(x[:need], dtype=np.complex64))
    out = []
    if sig.fdma:
        # one batched call: the channels are Doppler groups of the same (channel-independent) replica
        set_replicas(eng, sig, [None])
        f = np.concatenate([-(sig.carrier_step * chan + bins) / sig.fs for chan in keys])   # acquire-glonass-l1.py:28
        metric, lag, dbin = eng.search_grouped(f, len(bins), sig.n, B, sig.normalize, lag_limit)
        return [_finish(sig, L, bins, metric[g, 0], lag[g, 0], dbin[g, 0]) for g in range(len(keys))]
    set_replicas(eng, sig, keys)
    f = -bins / sig.fs                                              # acquire-gps-l1.py:28
    metric, lag, dbin = eng.search(f, sig.n, B, sig.normalize, lag_limit)
    return [_finish(sig, L, bins, metric[i], lag[i], dbin[i]) for i in range(len(keys))]


def search(signal, x, key, doppler_search, ms, engine=None):
    """Reference signature: search(x, prn, doppler_search, ms) of acquire-<signal>.py."""
    return acquire(signal, x, [key], doppler_search, ms, engine=engine)[0]


def format_result(signal, key, result):
    sig = SIGNALS[signal]
    metric, code, doppler = result
    return sig.fmt % (key, doppler, metric, code)
