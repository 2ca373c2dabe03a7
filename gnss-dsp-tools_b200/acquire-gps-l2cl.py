#!/usr/bin/env python
"""acquire-gps-l2cl.py — drop-in for the GNSS-DSP-tools script of the same name: same command
line, same preprocessing, same output line; the 75 code-phase hypotheses run on the GPU
correlator bank (gnsstools.acquire_serial, gnssacq_correlate_bank) instead of a Python loop."""

import optparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gnsstools import acquire_serial, io, nco     # noqa: E402

fs = None


def search(x, prn, doppler, l2cm_code_phase, ms):
    """Reference signature (acquire-gps-l2cl.py:18): returns (metric, k)."""
    return acquire_serial.search_l2cl(x, prn, doppler, l2cm_code_phase, ms, fs)


def main(argv=None):
    global fs
    parser = optparse.OptionParser(usage="""acquire-gps-l2cl.py [options] input_filename sample_rate carrier_offset prn doppler l2cm_code_phase

Acquire the GPS L2CL code phase given the L2CM acquisition result of the same PRN.

  input_filename    i/q interleaved, 8 bit signed
  sample_rate       Hz
  carrier_offset    offset to the L2C carrier in Hz
  prn, doppler, l2cm_code_phase   as printed by acquire-gps-l2cm.py""")
    parser.disable_interspersed_args()
    parser.add_option("--time", type="int", default=40, help="integration time in milliseconds (default %default)")
    options, args = parser.parse_args(argv)
    filename, fs, coffset = args[0], float(args[1]), float(args[2])
    prn, doppler, l2cm_code_phase = int(args[3]), float(args[4]), float(args[5])
    ms = options.time
    n = int(fs * 0.001 * (ms + 5))                       # acquire-gps-l2cl.py:70-73
    with open(filename, "rb") as fp:
        x = io.get_samples_complex(fp, n)
    nco.mix(x, -coffset / fs, 0)
    metric, k = search(x, prn, doppler, l2cm_code_phase, ms)
    print('%f %f' % (10230 * k + l2cm_code_phase, metric))


if __name__ == '__main__':
    main()
