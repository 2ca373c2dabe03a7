"""Kernel index logic on the CPU: the CUDA sources compiled against the host shim
(tests/cuda_emu) and driven through the same C ABI, compared with the oracle.
This is a test harness, not a product path — see tests/cuda_emu/cuda_emu.h."""
import numpy as np
import pytest

import emu_util
from oracle import acq_oracle as orc
import gnsstools.gps.ca as ca


@pytest.fixture(scope='module')
def eng():
    e = emu_util.emu_engine()
    yield e
    e.close()


def _case(eng, n, pad, boc, normalize, blocks, grid, fs, nprn=2, seed=0, lag_limit=None):
    rng = np.random.default_rng(seed)
    nx = n * (blocks + 2)
    x = rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)
    t = np.arange(nx)
    chips = [ca.ca_code(p) for p in range(1, nprn + 1)]
    x += 3 * orc.resample_code(chips[0], 300.25, 0, 1023.0 / n, nx) * np.exp(2j * np.pi * (grid[0] + 1.3 * grid[2]) * t / fs)
    x = x.astype(np.complex64)
    x64 = x.astype(np.complex128)
    eng.set_signal(x)
    eng.set_replicas(np.array([orc.replica(c, n, pad, boc) for c in chips]))
    f = -orc.doppler_bins(grid) / fs
    m, l, d, q = eng.search(f, n, blocks, normalize, n_lags=lag_limit or 0, dump=True)
    for i, c in enumerate(chips):
        ref, (idx, dbin, qg) = orc.search(x64, c, fs, n, grid, blocks, pad=pad, boc=boc, normalize=normalize,
                                          return_grid=True, lag_limit=lag_limit)
        assert int(l[i]) == idx and int(d[i]) == dbin
        assert abs(m[i] - ref[0]) <= 1e-4 * ref[0]           # north-star tolerance on the metric
        assert np.max(np.abs(q[i] - qg)) <= 1e-5 * np.max(qg)
    return eng.plan_info()


def test_mid_pow2_variant_a(eng):
    info = _case(eng, 4096, False, False, True, 2, (-1000, 1000, 500), 4.096e6)
    assert not info['large'] and info['N1'] * info['N2'] == 4096


def test_mid_mixed_radix_3_5_7_11(eng):
    _case(eng, 2310, False, False, True, 2, (-1000, 1000, 500), 2.31e6)


def test_mid_radix_31_boc(eng):
    _case(eng, 4092, False, True, False, 1, (-1000, 1000, 500), 4.092e6)


def test_mid_padded_three_blocks(eng):
    _case(eng, 2500, True, False, False, 3, (-1000, 1000, 500), 2.5e6)


def test_mid_radix_13(eng):
    _case(eng, 13 * 16 * 5, False, False, False, 1, (-500, 500, 500), 1.04e6, nprn=1)


@pytest.mark.slow
def test_large_pow2_padded(eng):
    info = _case(eng, 8192, True, False, False, 2, (-400, 400, 400), 8.192e6, nprn=1)
    assert info['large']


@pytest.mark.slow
def test_large_61380_like_l5(eng):
    info = _case(eng, 30690, True, False, False, 2, (-400, 400, 400), 30.69e6, nprn=1)
    assert info['large'] and info['N'] == 61380


def test_large_lag_limit(eng):
    _case(eng, 16368, False, False, True, 1, (-250, 250, 250), 16.368e6, nprn=1, lag_limit=1637)


def test_mix_matches_oracle(eng):
    rng = np.random.default_rng(5)
    for f, p in [(-9334875.0 / 69984000.0, 0.0), (0.01234, 0.25), (-0.49, 0.9)]:
        a = (rng.integers(-127, 128, 70001) + 1j * rng.integers(-127, 128, 70001)).astype(np.complex64)
        b = a.copy()
        eng.mix(a, f, p)
        orc.mix(b, f, p)
        assert np.array_equal(a, b)


def test_all_zero_input_selects_nothing(eng):
    n = 1024
    eng.set_signal(np.zeros(2 * n, np.complex64))
    eng.set_replicas(orc.replica(ca.ca_code(3), n, False, False)[None, :])
    m, l, d = eng.search(np.array([0.0, 1e-4]), n, 1, False)
    assert d[0] == -1 and m[0] == 0.0          # reference returns (0,0,0): strict '>' from 0


def test_errors(eng):
    eng.set_signal(np.zeros(100, np.complex64))
    eng.set_replicas(np.ones((1, 64), np.float32))
    with pytest.raises(ValueError):
        eng.search(np.array([0.0]), 64, 2, False)          # capture too short
    with pytest.raises(ValueError):
        eng.set_replicas(np.ones((1, 2 * 37), np.float32))  # prime factor 37 unsupported
