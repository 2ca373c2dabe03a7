"""Synthetic int8 IQ recordings for the command-line tests (seeded, regenerated on demand)."""
import numpy as np

from gnsstools import acquire
from gnsstools._codegen import resample
from gnsstools import nco

# name -> (script, fs_file, coffset, options, ms, planted (key, doppler, phase_chips, amp), seed)
CLI_CASES = {
    'config1-gps-l1': ('gps-l1', 4092000.0, 0.0, ['--prn', '1-4', '--doppler-search', '-5000,5000,500', '--time', '1'],
                       1, (1, 1500.0, 300.25, 4.0), 1234),
    'gps-l1-offset': ('gps-l1', 8184000.0, -1000000.0, ['--prn', '3,9', '--doppler-search', '-2000,2000,250', '--time', '3'],
                      3, (9, -750.0, 811.5, 3.0), 21),
    'glonass-l1': ('glonass-l1', 20000000.0, 250000.0, ['--channel', '-1:1', '--doppler-search', '-1000,1000,250', '--time', '2'],
                   2, (1, 250.0, 100.5, 3.0), 22),
    'galileo-e1b': ('galileo-e1b', 10000000.0, -500000.0, ['--prn', '11,12', '--doppler-search', '-200,200,50', '--time', '8'],
                    8, (11, 50.0, 2000.25, 3.0), 23),
    'gps-l5i': ('gps-l5i', 25000000.0, 0.0, ['--prn', '1-2', '--doppler-search', '-400,400,200', '--time', '2'],
                2, (2, 200.0, 5000.5, 3.0), 24),
    'gps-l1cd': ('gps-l1cd', 10000000.0, 100000.0, ['--prn', '4,5', '--doppler-search', '-40,40,20', '--time', '10'],
                 10, (5, 20.0, 7000.0, 3.0), 25),
    'beidou-b2bi-default-prns': ('beidou-b2bi', 31000000.0, 0.0, ['--doppler-search', '-200,200,200', '--time', '1'],
                                 1, (20, 0.0, 1234.5, 4.0), 26),
}


def recording(case):
    """int8 interleaved I/Q bytes of (ms+5) ms at the file rate, one planted satellite."""
    script, fs, coffset, _, ms, (key, doppler, phase, amp), seed = CLI_CASES[case]
    sig = acquire.SIGNALS[script]
    mod = acquire.code_module(sig)
    rng = np.random.default_rng(seed)
    n = int(fs * 0.001 * (ms + 5))
    t = np.arange(n)
    fn = getattr(mod, sig.module.split('.')[-1] + '_code')
    chips = np.asarray(fn() if sig.fdma else fn(key), dtype=np.float64)
    incr = mod.chip_rate / fs
    c = resample(chips, phase, 0, incr, n)
    if sig.boc:
        c = c * nco.boc11(phase, 0, incr, n)
    fc = coffset + doppler + (sig.carrier_step * key if sig.fdma else 0.0)
    x = amp * c * np.exp(2j * np.pi * fc * t / fs) + rng.normal(0, 8, n) + 1j * rng.normal(0, 8, n)
    iq = np.empty(2 * n, dtype=np.int8)
    iq[0::2] = np.clip(np.round(x.real), -127, 127)
    iq[1::2] = np.clip(np.round(x.imag), -127, 127)
    return iq.tobytes()


def command(case, path):
    script, fs, coffset, opts, _, _, _ = CLI_CASES[case]
    return script, opts + [path, repr(fs), repr(coffset)]


# Serial long-code acquisitions (acquire-gps-l2cl.py, acquire-glonass-l{1,2}-p.py):
# name -> (script, fs_file, coffset, options, ms, key (prn / channel), doppler, coarse phase, true k, amp, seed)
SERIAL_CASES = {
    'gps-l2cl': ('gps-l2cl', 2400000.0, -100000.0, ['--time', '40'], 40, 3, 431.0, 8317.2, 17, 2.0, 31),
    'glonass-l1-p': ('glonass-l1-p', 6000000.0, 250000.0, ['--time', '8'], 8, -2, 310.0, 278.6, 421, 2.0, 32),
    'glonass-l2-p': ('glonass-l2-p', 6000000.0, -125000.0, ['--time', '12'], 12, 3, -220.0, 33.4, 77, 2.0, 33),
}


def recording_serial(case):
    """int8 interleaved I/Q bytes of (ms+5) ms with the long code planted at hypothesis `true k`."""
    script, fs, coffset, _, ms, key, doppler, phase, k_true, amp, seed = SERIAL_CASES[case]
    rng = np.random.default_rng(seed)
    n = int(fs * 0.001 * (ms + 5))
    t = np.arange(n)
    if script == 'gps-l2cl':
        import gnsstools.gps.l2cl as l2cl
        c = resample(l2cl.l2cl_code(key), k_true * 10230 + phase, 0, l2cl.chip_rate / fs, n)
        fc = coffset + doppler
    else:
        import gnsstools.glonass.p as p
        c = resample(p.p_code(), 5110 * k_true + 10 * phase, 0, 5110000.0 / fs, n)
        fc = coffset + doppler + (562500 if script == 'glonass-l1-p' else 437500) * key
    x = amp * c * np.exp(2j * np.pi * fc * t / fs) + rng.normal(0, 8, n) + 1j * rng.normal(0, 8, n)
    iq = np.empty(2 * n, dtype=np.int8)
    iq[0::2] = np.clip(np.round(x.real), -127, 127)
    iq[1::2] = np.clip(np.round(x.imag), -127, 127)
    return iq.tobytes()


def command_serial(case, path):
    script, fs, coffset, opts, _, key, doppler, phase, _, _, _ = SERIAL_CASES[case]
    return script, opts + [path, repr(fs), repr(coffset), str(key), repr(doppler), repr(phase)]
