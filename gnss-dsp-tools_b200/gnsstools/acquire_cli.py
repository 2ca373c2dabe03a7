"""Command-line front end shared by the acquire-*.py scripts.

Same options, positional arguments, preprocessing and output lines as the reference scripts
(acquire-gps-l1.py:46-111 and siblings; per-script constants in acquire.SIGNALS):

    acquire-<signal>.py [--prn L | --channel L] [--doppler-search MIN,MAX,INCR] [--time MS] FILE FS COFFSET

read (ms+5) ms of int8 IQ -> nco.mix by -coffset/fs (GPU) -> 161-tap Hann FIR, zero-phase
(scipy filtfilt) -> linear-interpolation resample to the script's internal rate -> batched GPU
search over all PRNs/channels -> one text line per PRN. The reference's mp.Pool fan-out
(acquire-gps-l1.py:105-108) is replaced by the single batched call.
"""

import optparse
import sys

import numpy as np

from . import acquire as acq
from . import io, util
from . import _native


def preprocess(sig, raw, fs, coffset, ms_pad, engine=None):
    """Raw int8 I/Q at the file rate -> capture at the script's internal rate, resident on the
    GPU (acquire-gps-l1.py:85-96: nco.mix, firwin/filtfilt, np.interp). Only the 161 filter
    taps are computed on the host. Returns the number of samples produced."""
    import scipy.signal
    eng = engine if engine is not None else _native.default_engine()
    per_ms = int(round(sig.fs * 0.001))
    fsr = sig.fs / fs
    h = scipy.signal.firwin(161, sig.cutoff / (fs / 2), window='hann')
    n_out = ms_pad * per_ms
    eng.preprocess(raw, -coffset / fs, 0, h, (1 / fsr), n_out)
    return n_out


def preprocess_host(sig, x, fs, coffset, ms_pad, engine=None):
    """The same front end with scipy/numpy on the host, as the reference runs it (only nco.mix
    on the GPU); kept as the cross-check for preprocess()."""
    import scipy.signal
    eng = engine if engine is not None else _native.default_engine()
    eng.mix(x, -coffset / fs, 0)                                    # nco.mix(x,-coffset/fs,0)
    per_ms = int(round(sig.fs * 0.001))
    fsr = sig.fs / fs
    h = scipy.signal.firwin(161, sig.cutoff / (fs / 2), window='hann')
    x = scipy.signal.filtfilt(h, [1], x)
    t = (1 / fsr) * np.arange(ms_pad * per_ms)
    grid = np.arange(len(x))
    xr = np.interp(t, grid, np.real(x))
    xi = np.interp(t, grid, np.imag(x))
    return xr + (1j) * xi


def build_parser(name, sig):
    what = 'channels' if sig.fdma else 'PRNs'
    parser = optparse.OptionParser(usage="""acquire-%s.py [options] input_filename sample_rate carrier_offset

FFT acquisition search for %s on the GPU (B200). Command line and output lines follow the
GNSS-DSP-tools script of the same name.

Arguments:
  input_filename    input data file, i/q interleaved, 8 bit signed
  sample_rate       sampling rate in Hz
  carrier_offset    offset to the signal's carrier in Hz (positive or negative)""" % (name, name))
    parser.disable_interspersed_args()
    if sig.fdma:
        parser.add_option("--channel", default=sig.prns, help="channels to search, e.g. -6,-4,-1:2,7 (default %default)")
    else:
        parser.add_option("--prn", default=sig.prns, help="PRNs to search, e.g. 1,3,7-14,31 (default %default)")
    parser.add_option("--doppler-search", metavar="MIN,MAX,INCR", default=sig.doppler,
                      help="Doppler search grid: min,max,increment (default %default)")
    parser.add_option("--time", type="int", default=sig.time, help="integration time in milliseconds (default %default)")
    return parser


def default_keys(name, sig):
    """--prn "" in the B2b scripts means every PRN the ICD defines (acquire-beidou-b2bi.py:71)."""
    from . import _codegen
    return sorted(_codegen.memory_codes(sig.module).keys())


def main(name, argv=None, out=None, engine=None):
    sig = acq.SIGNALS[name]
    out = out if out is not None else sys.stdout
    (options, args) = build_parser(name, sig).parse_args(argv)
    filename = args[0]
    fs = float(args[1])
    coffset = float(args[2])
    if sig.fdma:
        keys = util.parse_list_ranges(options.channel, sep=':')
    elif not options.prn:
        keys = default_keys(name, sig)
    else:
        keys = util.parse_list_ranges(options.prn)
    doppler_search = util.parse_list_floats(options.doppler_search)
    ms = options.time

    ms_pad = ms + 5
    n = int(fs * 0.001 * ms_pad)
    with open(filename, "rb") as fp:
        raw = fp.read(2 * n)
    if len(raw) != 2 * n:
        # the reference gets None from io.get_samples_complex and dies with a TypeError inside nco.mix
        raise TypeError('short read: %s holds %d of the %d bytes needed' % (filename, len(raw), 2 * n))
    preprocess(sig, raw, fs, coffset, ms_pad, engine=engine)

    results = acq.acquire(name, None, keys, doppler_search, ms, engine=engine)
    for key, r in zip(keys, results):
        out.write(acq.format_result(name, key, r) + '\n')
    return results
