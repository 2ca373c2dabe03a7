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
    # 279 = 31*9 and 220 = 11*20 are coprime schedules: both tile transforms run twiddle-free (prime-factor),
    # and gcd(279, 220) = 1: the four-step split is coprime too (no twiddle pass)
    assert eng.kernel_variant() & 56 == 56
    # the generic Cooley-Tukey kernels must agree with the prime-factor ones
    eng.set_option('specialized_kernels', 0)
    try:
        _case(eng, 30690, True, False, False, 2, (-400, 400, 400), 30.69e6, nprn=1)
        assert eng.kernel_variant() == 0
    finally:
        eng.set_option('specialized_kernels', 1)


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
        eng.set_replicas(np.ones((1, 2 * 300007), np.float32))  # not plannable and too long to embed (> 524288)


@pytest.mark.parametrize('n,pad,boc,normalize,blocks,lag_limit', [(646, False, False, True, 2, None),       # 2*17*19 -> 2048 (one CTA)
                                                                  (4522, True, False, False, 2, None),        # 2n = 4*7*17*19 -> 32768 (two kernels)
                                                                  (4097, False, True, False, 1, 1000)])       # 17*241 -> 16384
def test_any_length_runs_embedded_in_a_power_of_two(eng, n, pad, boc, normalize, blocks, lag_limit):
    """Lengths the planner cannot factor (the reference takes any N through fftpack,
    acquire-gps-l5i.py:24,32) run embedded in the next power of two >= 2N-1: periodically extended
    replica, zero-padded blocks, lags 0..N-1 — same q grid, same indices, same mean."""
    info = _case(eng, n, pad, boc, normalize, blocks, (-400, 400, 400), n * 1000.0, nprn=1, lag_limit=lag_limit)
    N = 2 * n if pad else n
    assert info['N'] >= 2 * N - 1 and info['N'] & (info['N'] - 1) == 0 and eng.kernel_variant() == 128
    _case(eng, 4096, False, False, True, 1, (-1000, 1000, 500), 4.096e6)      # back to a plannable length
    assert eng.kernel_variant() == 4


# --------------------------------------------------------------------------- replica builder / correlator bank
@pytest.mark.parametrize('module,key,n,pad,boc', [('gps.ca', 5, 2046, False, False), ('galileo.e1b', 11, 4092, True, True),
                                                  ('gps.l5i', 3, 2500, True, False), ('glonass.ca', None, 1000, False, False)])
def test_replica_builder_equals_host_replica(eng, module, key, n, pad, boc):
    """gnssacq_set_replicas_from_chips builds exactly the samples acquire.replica() builds on the host
    (reference <sig>.code x nco.boc11, zero half): identical replicas -> bit-identical q grids."""
    from gnsstools import acquire
    sig = acquire.Signal(module, 1.0e6, n, lambda ms: ms, pad=pad, boc=boc, fdma=key is None)
    rng = np.random.default_rng(3)
    x = (rng.normal(0, 8, 3 * n) + 1j * rng.normal(0, 8, 3 * n)).astype(np.complex64)
    f = np.array([0.0, 1.5e-4])
    eng.set_signal(x)
    eng.set_replicas(acquire.replica(sig, key)[None, :])
    a = eng.search(f, n, 1, False, dump=True)
    acquire.set_replicas(eng, sig, [key])
    b = eng.search(f, n, 1, False, dump=True)
    assert np.array_equal(a[3], b[3]) and a[1][0] == b[1][0] and a[0][0] == b[0][0]


def test_correlator_bank_matches_numpy(eng):
    rng = np.random.default_rng(4)
    L, n, B, H = 5000, 3000, 3, 7
    chips = rng.integers(0, 2, L).astype(np.int8)
    x = (rng.normal(0, 8, B * n) + 1j * rng.normal(0, 8, B * n)).astype(np.complex64)
    f, incr = -1.234e-3, 0.731
    base = rng.uniform(0, L, (H, B))
    base[0, 0] = L - 0.25                      # wraps inside the block
    eng.set_signal(x)
    got = eng.correlate_bank(chips, f, n, B, n, base, incr)
    w = orc.nco(f, 0, n)
    for h in range(H):
        for b in range(B):
            c = orc.resample_code(chips, 0, base[h, b], incr, n)
            want = np.sum(x[b * n:(b + 1) * n].astype(np.complex128) * c * w)
            assert abs(got[h, b] - want) <= 1e-5 * np.sum(np.abs(x[b * n:(b + 1) * n])), (h, b)


def test_serial_searches_match_oracle(eng):
    """acquire-gps-l2cl.py / acquire-glonass-l1-p.py search() through the correlator bank."""
    from gnsstools import acquire_serial
    import gnsstools.gps.l2cl as l2cl
    import gnsstools.glonass.p as gp
    rng = np.random.default_rng(6)
    fs, ms, prn, doppler, cm = 0.3e6, 40, 3, 431.0, 8317.2
    nx = int(fs * 0.001 * (ms + 5))
    t = np.arange(nx)
    x = rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)
    x += 3.0 * orc.resample_code(l2cl.l2cl_code(prn), 4 * 10230 + cm, 0, 511500.0 / fs, nx) * np.exp(2j * np.pi * doppler * t / fs)
    x = x.astype(np.complex64)
    got = acquire_serial.search_l2cl(x, prn, doppler, cm, ms, fs, engine=eng, hypotheses=9)
    want = orc.search_l2cl(x, l2cl.l2cl_code(prn), fs, doppler, cm, ms, hypotheses=9)
    assert got[1] == want[1] == 4 and abs(got[0] - want[0]) <= 1e-4 * want[0]
    fs, ms, chan, doppler, ca = 1.0e6, 8, -2, 310.0, 278.6
    nx = int(fs * 0.001 * (ms + 5))
    t = np.arange(nx)
    x = rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)
    x += 3.0 * orc.resample_code(gp.p_code(), 5110 * 6 + 10 * ca, 0, 5110000.0 / fs, nx) * np.exp(2j * np.pi * (562500 * chan + doppler) * t / fs)
    x = x.astype(np.complex64)
    got = acquire_serial.search_glonass_p(x, chan, doppler, ca, ms, fs, 562500, engine=eng, hypotheses=10)
    want = orc.search_glonass_p(x, gp.p_code(), fs, 562500, chan, doppler, ca, ms, hypotheses=10)
    assert got[1] == want[1] == 6 and abs(got[0] - want[0]) <= 1e-4 * want[0]


@pytest.mark.slow
def test_large_163680_coprime_split_341x480(eng):
    """BASELINE config 2's transform as a coprime (Good-Thomas) four-step: 163680 = 341 x 480, no
    twiddle pass; 341 = 31*11 and 480 = 15*32 are themselves twiddle-free two-stage schedules."""
    info = _case(eng, 163680, False, False, True, 1, (-250, 250, 250), 16.368e6, nprn=1, lag_limit=16368)
    assert info['N1'] == 341 and info['N2'] == 480 and eng.kernel_variant() == 123      # 64: copy-engine-fed pair


@pytest.mark.slow
def test_large_163680_register_loading_kernels_on_the_coprime_split(eng):
    """v3 = 0: the same plan through the register-loading kernels (kernels_small.cuh)."""
    eng.set_option('v3', 0)
    try:
        info = _case(eng, 163680, False, False, True, 1, (-250, 250, 250), 16.368e6, nprn=1, lag_limit=16368)
        assert info['N1'] == 341 and eng.kernel_variant() == 59
    finally:
        eng.set_option('v3', 1)


@pytest.mark.slow
@pytest.mark.parametrize('rows,cols,rc,g', [(0, 0, 2, 2), (1, 1, 1, 3), (2, 2, 3, 1), (3, 0, 2, 4), (4, 0, 2, 5), (4, 1, 3, 1), (0, 3, 2, 2), (0, 4, 3, 5), (5, 0, 2, 2), (5, 1, 3, 5)])
def test_v3_tile_shapes_and_chunk_edges_163680(eng, rows, cols, rc, g):
    """Every instantiated tile shape of the copy-engine-fed pair, with chunk shapes that leave ragged
    edges (3 replicas in chunks of rc, 5 Doppler bins in groups of g)."""
    for k, v in (('v3_rows', rows), ('v3_cols', cols), ('v3_rc', rc), ('v3_g', g)):
        eng.set_option(k, v)
    try:
        _case(eng, 163680, False, False, True, 1, (-500, 750, 250), 16.368e6, nprn=3, lag_limit=16368)
        assert eng.kernel_variant() & (64 | 256) == 64
    finally:
        for k in ('v3_rows', 'v3_cols', 'v3_rc', 'v3_g'):
            eng.set_option(k, 0)


@pytest.mark.slow
@pytest.mark.parametrize('rc,g,sets,tpt', [(2, 2, 2, 1), (1, 3, 3, 7), (3, 4, 2, 60)])
def test_fused_kernel_group_shapes_163680(eng, rc, g, sets, tpt):
    """The fused persistent kernel with group shapes that leave ragged edges (3 replicas, 5 bins)
    and with the shortest scratch ring (every rows group waits for the columns tasks two groups back)."""
    for k, v in (('fused', 1), ('fused_rc', rc), ('fused_g', g), ('fused_sets', sets), ('fused_tpt', tpt)):
        eng.set_option(k, v)
    try:
        _case(eng, 163680, False, False, True, 1, (-500, 750, 250), 16.368e6, nprn=3, lag_limit=16368)
        assert eng.kernel_variant() & 256
    finally:
        eng.set_option('fused_rc', 0)
        eng.set_option('fused_g', 0)
        eng.set_option('fused_sets', 3)
        eng.set_option('fused_tpt', 0)
        eng.set_option('fused', 0)


@pytest.mark.slow
@pytest.mark.parametrize('rows,cols,blocks', [(0, 0, 3), (1, 1, 2), (0, 3, 3), (0, 6, 3), (1, 7, 2)])   # 279 x 220: tile shapes; 3 = single-slot columns kernel; 6, 7 = non-coherent sums in registers
def test_v3_non_coherent_blocks_61380(eng, rows, cols, blocks):
    """279 x 220 with several non-coherent blocks: the (Doppler, block) list walked by the rows
    kernel (split over grid.z), q accumulated in shared memory by the columns kernel."""
    eng.set_option('v3_rows', rows)
    eng.set_option('v3_cols', cols)
    try:
        for fused in (0, 1):
            eng.set_option('fused', fused)
            _case(eng, 30690, True, False, False, blocks, (-400, 400, 200), 30.69e6, nprn=2)
            assert eng.kernel_variant() & 64 and bool(eng.kernel_variant() & 256) == bool(fused)
    finally:
        eng.set_option('v3_rows', 0)
        eng.set_option('v3_cols', 0)
        eng.set_option('fused', 0)


@pytest.mark.slow
def test_large_163680_prime_factor_31x12(eng):
    """The Cooley-Tukey split of the same length (gt_split = 0): 372 x 440 with 372 = 31*12 and
    440 = 11*5*8, prime-factor tile transforms, four-step twiddles between them."""
    eng.set_option('gt_split', 0)
    try:
        info = _case(eng, 163680, False, False, True, 1, (-250, 250, 250), 16.368e6, nprn=1, lag_limit=16368)
        assert info['N1'] == 372 and info['N2'] == 440 and eng.kernel_variant() == 27
    finally:
        eng.set_option('gt_split', 1)


@pytest.mark.slow
def test_coprime_split_generic_kernels_any_length(eng):
    """A forced coprime split of a length without specialised kernels runs the generic kernels in
    Good-Thomas form: 2 * 5 * 7 * 9 * 13 = 8190 -> 90 x 91... too short for a large plan, so 16380 = 130 x 126?
    gcd(130, 126) = 2; 16380 = 4 * 9 * 5 * 7 * 13 = 180 x 91 (coprime)."""
    eng.set_option('split_n1', 91)
    try:
        info = _case(eng, 16380, False, False, False, 2, (-300, 300, 300), 16.38e6, nprn=1)
        assert info['N1'] == 91 and info['N2'] == 180 and eng.kernel_variant() & 32
    finally:
        eng.set_option('split_n1', 0)


@pytest.mark.slow
def test_large_mixed_prime_factor_and_generic(eng):
    """32736 = 186 x 176: the columns transform (31*6) is a specialised prime-factor one, the rows
    transform (11*16) has no specialisation and runs the generic Cooley-Tukey kernels."""
    info = _case(eng, 32736, False, False, True, 1, (-250, 250, 250), 16.368e6, nprn=1)
    assert info['N1'] == 186 and info['N2'] == 176 and eng.kernel_variant() == (2 | 8)


def test_grouped_search_equals_one_search_per_group(eng):
    """gnssacq_search_grouped (the FDMA channel loop of acquire-glonass-l1.py:60-69 as one batch):
    every group's records equal those of a separate search over that group's Doppler list."""
    rng = np.random.default_rng(1)
    n = 2048
    x = (rng.normal(0, 8, 3 * n) + 1j * rng.normal(0, 8, 3 * n)).astype(np.complex64)
    eng.set_signal(x)
    eng.set_replicas(np.array([orc.replica(ca.ca_code(p), n, False, False) for p in (1, 2)]))
    bins = np.arange(-3, 3) * 1e-4
    groups = [bins + g * 0.01 for g in range(4)]
    M, L, Dd = eng.search_grouped(np.concatenate(groups), len(bins), n, 2, True)
    assert M.shape == (4, 2)
    for g, f in enumerate(groups):
        m, l, d = eng.search(f, n, 2, True)
        assert np.array_equal(m, M[g]) and np.array_equal(l, L[g]) and np.array_equal(d, Dd[g])
    with pytest.raises(ValueError):
        eng.search_grouped(np.zeros(7), 3, n, 2, True)          # not a whole number of groups


@pytest.mark.slow
def test_30690_coprime_split_341x90_with_ragged_column_tiles(eng):
    """30690 = 341 x 90 (E6, Xona X5): 90 = 9 * 10 has rows of 10-element groups, which 8-column tiles do not
    divide — the columns kernel walks the flat padded row and masks what lies behind it."""
    info = _case(eng, 15345, True, False, False, 2, (-400, 400, 200), 15.345e6, nprn=2)
    assert (info['N1'], info['N2']) == (341, 90) and eng.kernel_variant() == 123
    _case(eng, 30690, False, False, True, 2, (-400, 400, 400), 30.69e6, nprn=1)
    eng.set_option('v3', 0)
    try:
        _case(eng, 15345, True, False, False, 2, (-400, 400, 200), 15.345e6, nprn=2)
        assert eng.kernel_variant() == 59
    finally:
        eng.set_option('v3', 1)


def test_short_last_doppler_chunk_uses_a_lane_buffer(eng):
    """A search whose Doppler list splits into a long chunk (alternating over the lanes' scratch buffers) and a
    short last one (single stream): the short chunk must find a scratch buffer although the handle's own is
    not allocated for lane searches."""
    eng.set_option('xchunk_mb', 1)                  # 16384-point spectra: 8 Doppler bins per chunk
    eng.set_option('units_per_chunk', 4)
    try:
        _case(eng, 8192, True, False, False, 1, (-800, 1000, 200), 8.192e6, nprn=3)     # 9 bins: chunks of 8 (24 units, lanes) + 1 (3 units)
    finally:
        eng.set_option('xchunk_mb', 512)
        eng.set_option('units_per_chunk', 0)
