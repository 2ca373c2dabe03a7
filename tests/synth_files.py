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
    # the other 22 FFT scripts, one small case each (format strings, block-count rules, defaults)
    'xona-x1': ('xona-x1', 8000000.0, 50000.0, ['--doppler-search', '-1000,1000,250', '--time', '2'],
                2, (0, -500.0, 411.25, 4.0), 40),
    'xona-x5p': ('xona-x5p', 31000000.0, 0.0, ['--prn', '0', '--doppler-search', '-400,400,200', '--time', '2'],
                 2, (0, 200.0, 6001.5, 4.0), 41),
    'glonass-l2': ('glonass-l2', 20000000.0, -150000.0, ['--channel', '2:3', '--doppler-search', '-500,500,250', '--time', '2'],
                   2, (3, -250.0, 77.5, 3.0), 42),
    'gps-l1cp': ('gps-l1cp', 10000000.0, 0.0, ['--prn', '7,8', '--doppler-search', '-40,40,20', '--time', '10'],
                 10, (7, -20.0, 3100.5, 3.0), 43),
    'beidou-b1cd': ('beidou-b1cd', 10000000.0, -50000.0, ['--prn', '19,20', '--doppler-search', '-40,40,20', '--time', '10'],
                    10, (20, 0.0, 9000.25, 3.0), 44),
    'beidou-b1cp': ('beidou-b1cp', 10000000.0, 0.0, ['--prn', '33', '--doppler-search', '-40,40,20', '--time', '20'],
                    20, (33, 20.0, 123.0, 3.0), 45),
    'galileo-e1c': ('galileo-e1c', 10000000.0, 0.0, ['--prn', '5,6', '--doppler-search', '-100,100,50', '--time', '12'],
                    12, (6, -50.0, 1000.75, 3.0), 46),
    'beidou-b1i': ('beidou-b1i', 10000000.0, 100000.0, ['--prn', '1-3', '--doppler-search', '-400,400,200', '--time', '3'],
                   3, (2, 200.0, 1500.5, 3.0), 47),
    'beidou-b2i': ('beidou-b2i', 10000000.0, 0.0, ['--prn', '6,7', '--doppler-search', '-400,400,200', '--time', '2'],
                   2, (6, -200.0, 20.25, 3.0), 48),
    'gps-l2cm': ('gps-l2cm', 5000000.0, 0.0, ['--prn', '3,4', '--doppler-search', '-40,40,20', '--time', '40'],
                 40, (4, 20.0, 5115.5, 3.0), 49),
    'gps-l5q': ('gps-l5q', 31000000.0, 0.0, ['--prn', '10,11', '--doppler-search', '-400,400,200', '--time', '2'],
                2, (11, -200.0, 10000.5, 3.0), 50),
    'galileo-e5ai': ('galileo-e5ai', 31000000.0, 0.0, ['--prn', '1,2', '--doppler-search', '-400,400,200', '--time', '2'],
                     2, (1, 0.0, 42.5, 3.0), 51),
    'galileo-e5aq': ('galileo-e5aq', 31000000.0, 0.0, ['--prn', '12', '--doppler-search', '-400,400,200', '--time', '3'],
                     3, (12, 200.0, 9999.0, 3.0), 52),
    'galileo-e5bi': ('galileo-e5bi', 31000000.0, 250000.0, ['--prn', '24,25', '--doppler-search', '-400,400,200', '--time', '2'],
                     2, (25, -400.0, 3333.25, 3.0), 53),
    'galileo-e5bq': ('galileo-e5bq', 31000000.0, 0.0, ['--prn', '36', '--doppler-search', '-400,400,200', '--time', '2'],
                     2, (36, 0.0, 7777.75, 3.0), 54),
    'galileo-e6b': ('galileo-e6b', 16000000.0, 0.0, ['--prn', '8,9', '--doppler-search', '-400,400,200', '--time', '3'],
                    3, (9, 200.0, 2557.5, 3.0), 55),
    'galileo-e6c': ('galileo-e6c', 16000000.0, -80000.0, ['--prn', '30', '--doppler-search', '-400,400,200', '--time', '2'],
                    2, (30, -200.0, 5000.25, 3.0), 56),
    # B2a data: search() always integrates 80 blocks whatever --time says (acquire-beidou-b2ad.py:29)
    'beidou-b2ad': ('beidou-b2ad', 31000000.0, 0.0, ['--prn', '21,22', '--doppler-search', '-200,200,200', '--time', '80'],
                    80, (22, 0.0, 4321.5, 1.0), 57),
    'beidou-b2ap': ('beidou-b2ap', 31000000.0, 0.0, ['--prn', '9,10', '--doppler-search', '-400,400,200', '--time', '2'],
                    2, (9, 200.0, 8765.25, 3.0), 58),
    'beidou-b2bq': ('beidou-b2bq', 31000000.0, 0.0, ['--prn', '32-34', '--doppler-search', '-200,200,200', '--time', '2'],
                    2, (33, -200.0, 100.5, 3.0), 59),
    'beidou-b3i': ('beidou-b3i', 31000000.0, 0.0, ['--prn', '40,41', '--doppler-search', '-400,400,200', '--time', '2'],
                   2, (41, 0.0, 6000.0, 3.0), 60),
    'glonass-l3ocd': ('glonass-l3ocd', 31000000.0, 0.0, ['--prn', '0,1', '--doppler-search', '-400,400,200', '--time', '2'],
                      2, (0, -200.0, 512.5, 3.0), 61),
    'glonass-l3ocp': ('glonass-l3ocp', 31000000.0, 0.0, ['--prn', '62,63', '--doppler-search', '-400,400,200', '--time', '2'],
                      2, (63, 200.0, 10229.5, 3.0), 62),
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
