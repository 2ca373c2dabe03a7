#!/usr/bin/env python
"""acquire-glonass-l2-p.py — drop-in for the GNSS-DSP-tools script of the same name: same command
line, same preprocessing, same output line; the 1000 code-phase hypotheses run on the GPU
correlator bank (gnsstools.acquire_serial, gnssacq_correlate_bank) instead of a Python loop."""

import optparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gnsstools import acquire_serial, io, nco     # noqa: E402

fs = None
CARRIER_STEP = 437500          # FDMA channel spacing in Hz (acquire-glonass-l2-p.py:18)


def search(x, chan, doppler, ca_code_phase, ms):
    """Reference signature (acquire-glonass-l2-p.py:14): returns (metric, k)."""
    return acquire_serial.search_glonass_p(x, chan, doppler, ca_code_phase, ms, fs, CARRIER_STEP)


def main(argv=None):
    global fs
    parser = optparse.OptionParser(usage="""acquire-glonass-l2-p.py [options] input_filename sample_rate carrier_offset channel doppler ca_code_phase

Acquire the GLONASS L2-P code phase given the L2-CA acquisition result of the same RF channel.

  input_filename    i/q interleaved, 8 bit signed
  sample_rate       Hz
  carrier_offset    offset to the GLONASS L2 carrier (channel 0) in Hz
  channel, doppler, ca_code_phase   as printed by acquire-glonass-l2.py""")
    parser.disable_interspersed_args()
    parser.add_option("--time", type="int", default=80, help="integration time in milliseconds (default %default)")
    options, args = parser.parse_args(argv)
    filename, fs, coffset = args[0], float(args[1]), float(args[2])
    chan, doppler, ca_code_phase = int(args[3]), float(args[4]), float(args[5])
    ms = options.time
    n = int(fs * 0.001 * (ms + 5))                       # acquire-glonass-l2-p.py:77-80
    with open(filename, "rb") as fp:
        x = io.get_samples_complex(fp, n)
    nco.mix(x, -coffset / fs, 0)
    metric, k = search(x, chan, doppler, ca_code_phase, ms)
    print('%f %f' % (5110 * k + 10 * ca_code_phase, metric))


if __name__ == '__main__':
    main()
