"""Shared machinery for the PRN code generator modules.

Every generator module exposes the reference surface
``chip_rate, code_length, codes, <sig>_code(prn), code(prn, chips, frac, incr, n)``
(e.g. reference gnsstools/gps/ca.py:101-112). The resampling rule is common and lives here,
as do the shift-register, Weil-code and table-loading helpers. Per-PRN parameters and
memory codes are ICD constants kept as data under _data/ (tools/extract_code_tables.py).
"""

import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_data')
_tables = None
_memory = None


def resample(chip_bits, chips, frac, incr, n):
    """+-1.0 float64[n]: 1 - 2*c[floor((chips % L) + frac + incr*i) mod L]
    (reference gnsstools/gps/ca.py:106-112; identical body in every module)."""
    L = len(chip_bits)
    ph = (chips % L) + frac + incr * np.arange(n)
    k = np.mod(np.floor(ph).astype('int'), L)
    return 1.0 - 2.0 * chip_bits[k]


def icd_table(module, name):
    """{prn: value} for one per-PRN parameter table of `module` (e.g. 'gps.l5i', 'l5i_init')."""
    global _tables
    if _tables is None:
        with open(os.path.join(_DATA, 'icd_tables.json')) as f:
            _tables = json.load(f)
    return {int(k): (tuple(v) if isinstance(v, list) else v) for k, v in _tables[module][name].items()}


def _npz():
    global _memory
    if _memory is None:
        _memory = np.load(os.path.join(_DATA, 'memory_codes.npz'))
    return _memory


def memory_codes(module):
    """{prn: float64 0/1 chips} for a memory-code signal (Galileo E1/E6, BeiDou B2b I/Q, Xona)."""
    z = _npz()
    L = int(z[module + ':length'])
    bits = np.unpackbits(z[module + ':bits'], axis=1)[:, :L].astype(np.float64)
    return {int(p): bits[i] for i, p in enumerate(z[module + ':prns'])}


def secondary_table(module):
    """{prn: +-1 float64 secondary code} for signals whose secondary codes are tabulated per PRN."""
    z = _npz()
    L = int(z[module + ':sec_length'])
    bits = np.unpackbits(z[module + ':sec_bits'], axis=1)[:, :L].astype(np.float64)
    return {int(p): 1.0 - 2.0 * bits[i] for i, p in enumerate(z[module + ':sec_prns'])}


# ----------------------------------------------------------------------------- shift registers
# Register stages x[0..n-1] are held in an int with bit k = x[k]. One Fibonacci step is
# x <- [xor of tapped stages] + x[0:n-1], i.e. shift towards the high bit, feedback into bit 0.

def _parity(v):
    return bin(v).count('1') & 1


def mask_of(stages):
    m = 0
    for s in stages:
        m |= 1 << s
    return m


def bits_to_int(bits):
    """x[k] = bits[k] -> int with bit k = x[k] (accepts a list or a '0101' string)."""
    v = 0
    for k, b in enumerate(bits):
        if int(b):
            v |= 1 << k
    return v


def lfsr_states(nbits, taps, state, length, reset_after=None, reset_from=None, reset_to=None):
    """States of a Fibonacci LFSR before each of `length` steps, as an int64 array.

    reset_after=i: after producing sample i the register is reloaded with `reset_to`
    instead of being stepped (BeiDou B2 short cycle at i == 8189).
    reset_from=s: whenever the current state equals s, the next state is `reset_to`
    (GPS L5 XA / BeiDou B3I short cycles).
    """
    full = (1 << nbits) - 1
    tapmask = mask_of(taps)
    out = np.empty(length, dtype=np.int64)
    x = state
    for i in range(length):
        out[i] = x
        if reset_after is not None and i == reset_after:
            x = reset_to
        elif reset_from is not None and x == reset_from:
            x = reset_to
        else:
            x = ((x << 1) & full) | _parity(x & tapmask)
    return out


def stage(states, k):
    """Output sequence of register stage x[k] as float64 0/1."""
    return ((states >> k) & 1).astype(np.float64)


def lfsr_fibonacci(nbits, taps, state, length, out_tap=None):
    """Sequence of stage `out_tap` (default: last) of a Fibonacci LFSR started at `state`
    (int, list of bits x[0..], or bit string)."""
    if not isinstance(state, int):
        state = bits_to_int(state)
    return stage(lfsr_states(nbits, taps, state, length), nbits - 1 if out_tap is None else out_tap)


def lfsr_galois_lsb(poly, state, length):
    """Galois register stepping x <- (x >> 1) ^ (x & 1) * poly, output x & 1 (GPS L2C)."""
    out = np.empty(length)
    x = state
    for i in range(length):
        out[i] = x & 1
        x = (x >> 1) ^ (poly if x & 1 else 0)
    return out, x


# ----------------------------------------------------------------------------- Weil codes
def legendre_sequence(N):
    """L[i] = 1 when i is a non-zero quadratic residue mod the prime N, else 0."""
    L = np.zeros(N, dtype=np.int64)
    L[(np.arange(1, N, dtype=np.int64) ** 2) % N] = 1
    return L


def weil(L, w):
    """W[k] = L[k] xor L[(k + w) mod N]."""
    return L ^ np.roll(L, -w)


def weil_truncated(L, w, p, length):
    """c[n] = W[(n + p - 1) mod N], n < length (BeiDou B1C / B2a secondary construction)."""
    N = len(L)
    return weil(L, w)[(np.arange(length) + p - 1) % N].astype(np.float64)


def hex_to_bits(s, nbits):
    """First nbits of a hex string, MSB of each nibble first."""
    v = np.array([(int(s[i // 4], 16) >> (3 - (i % 4))) & 1 for i in range(nbits)])
    return v.astype(np.float64)


# ----------------------------------------------------------------------------- tracking correlators
# Out of the acquisition hot path (tracking is out of scope); kept so that code
# written against the reference modules still imports. Plain restatement of the reference's
# scalar loop (e.g. gnsstools/gps/ca.py:120-128), Numba-compiled when Numba is present.
try:
    from numba import jit as _jit
except Exception:                                   # pragma: no cover
    def _jit(**kwargs):
        return lambda f: f


@_jit(nopython=True)
def correlate_plain(x, chips, frac, incr, c, code_length):
    p = 0.0j
    cp = (chips + frac) % code_length
    for i in range(len(x)):
        p += x[i] * (1.0 - 2.0 * c[int(cp)])
        cp = (cp + incr) % code_length
    return p


@_jit(nopython=True)
def correlate_sub2(x, chips, frac, incr, c, code_length, sub):
    """Code times a 2-level half-chip pattern `sub` (BOC(1,1): [1,-1]; L2C RZ slot: [1,0])."""
    p = 0.0j
    cp = (chips + frac) % code_length
    bp = (2 * (chips + frac)) % 2
    for i in range(len(x)):
        p += x[i] * (1.0 - 2.0 * c[int(cp)]) * sub[int(bp)]
        cp = (cp + incr) % code_length
        bp = (bp + 2 * incr) % 2
    return p


@_jit(nopython=True)
def correlate_cboc(x, chips, frac, incr, c, code_length, boc11, a1, a6):
    """Galileo E1 CBOC: a1*BOC(1,1) + a6*BOC(6,1) (reference galileo/e1b.py:45-58)."""
    p = 0.0j
    cp = (chips + frac) % code_length
    bp = (2 * (chips + frac)) % 2
    bp6 = (12 * (chips + frac)) % 2
    for i in range(len(x)):
        cboc = a1 * boc11[int(bp)] + a6 * boc11[int(bp6)]
        p += x[i] * (1.0 - 2.0 * c[int(cp)]) * cboc
        cp = (cp + incr) % code_length
        bp = (bp + 2 * incr) % 2
        bp6 = (bp6 + 12 * incr) % 2
    return p


@_jit(nopython=True)
def correlate_tmboc(x, chips, frac, incr, c, code_length, boc11, pattern):
    """GPS L1Cp TMBOC: BOC(6,1) on the chips flagged in the 33-chip pattern
    (reference gps/l1cp.py:210-228)."""
    p = 0.0j
    cp = (chips + frac) % code_length
    bp = (2 * (chips + frac)) % 2
    bp6 = (12 * (chips + frac)) % 2
    u = int(cp % 33)
    for i in range(len(x)):
        boc = boc11[int(bp6)] if pattern[u] else boc11[int(bp)]
        p += x[i] * (1.0 - 2.0 * c[int(cp)]) * boc
        cp = (cp + incr) % code_length
        bp = (bp + 2 * incr) % 2
        bp6 = (bp6 + 12 * incr) % 2
        u = int(cp % 33)
    return p
