#!/usr/bin/env python
"""acquire-gps-l1cd.py — drop-in for the GNSS-DSP-tools script of the same name: same command
line, same preprocessing, same output lines; the FFT search runs on the GPU through
libgnssacq.so (gnsstools.acquire) instead of a per-PRN multiprocessing pool."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gnsstools import acquire, acquire_cli     # noqa: E402


def search(x, prn, doppler_search, ms):
    """Reference signature (acquire-gps-l1cd.py:18): returns (metric, code_chips, doppler_hz)."""
    return acquire.search('gps-l1cd', x, prn, doppler_search, ms)


if __name__ == '__main__':
    acquire_cli.main('gps-l1cd')
