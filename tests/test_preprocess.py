"""GPU capture front end (nco.mix -> filtfilt -> np.interp) against the scipy/numpy pipeline the
reference runs on the host (acquire-gps-l1.py:80-96)."""
import numpy as np
import pytest
import scipy.signal

from oracle import acq_oracle as orc


def reference_front_end(raw, fs, fs_int, cutoff, coffset, ms_pad):
    x = raw.astype(np.float32).view(np.complex64).copy()          # io.get_samples_complex
    orc.mix(x, -coffset / fs, 0)
    fsr = fs_int / fs
    h = scipy.signal.firwin(161, cutoff / (fs / 2), window='hann')
    y = scipy.signal.filtfilt(h, [1], x)
    n_out = ms_pad * int(round(fs_int * 0.001))
    t = (1 / fsr) * np.arange(n_out)
    g = np.arange(len(y))
    return np.interp(t, g, np.real(y)) + 1j * np.interp(t, g, np.imag(y)), h, (1 / fsr), n_out


def check(eng, fs, fs_int, cutoff, coffset, ms_pad, seed):
    rng = np.random.default_rng(seed)
    n = int(fs * 0.001 * ms_pad)
    raw = rng.integers(-127, 128, 2 * n).astype(np.int8)
    want, h, step, n_out = reference_front_end(raw, fs, fs_int, cutoff, coffset, ms_pad)
    got = eng.preprocess(raw, -coffset / fs, 0.0, h, step, n_out, return_c128=True)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
    # what the search consumes is identical
    assert np.array_equal(got.astype(np.complex64), want.astype(np.complex64))


def test_front_end_on_emulated_kernels():
    import emu_util
    eng = emu_util.emu_engine()
    check(eng, 5.0e6, 4.096e6, 1.5e6, -123456.0, 3, 1)            # downsample
    check(eng, 4.092e6, 4.096e6, 1.5e6, 0.0, 2, 2)                # config-1 rates: slight upsample, right-edge clamp
    with pytest.raises(ValueError):
        eng.preprocess(np.zeros(2 * 400, np.int8), 0.0, 0.0, np.ones(161), 1.0, 100)   # shorter than the padding
    eng.close()


@pytest.mark.gpu
def test_front_end_on_gpu():
    from gnsstools import _native
    eng = _native.Engine(0)
    check(eng, 69.984e6, 4.096e6, 1.5e6, -9334875.0, 25, 3)       # the reference's own recording rates
    check(eng, 25.0e6, 30.69e6, 12e6, 1000.0, 6, 4)
    eng.close()
