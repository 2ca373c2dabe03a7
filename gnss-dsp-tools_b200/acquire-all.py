#!/usr/bin/env python
"""acquire-all.py — the sweep of the reference's acquire-all.sh in one process on the GPU.

    acquire-all.py [--fs 69984000] [--time 80] L1_FILE L2_FILE L5_FILE DEST_DIR

The three files are the per-band int8 I/Q recordings that acquire-all.sh obtains from
`packet2wav_3ch 1|2|3`. One acq-<signal>.dat per entry of the reference's list is written to
DEST_DIR. Under torchrun (one process per GPU) the signals are sharded across the ranks."""

import optparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gnsstools import sweep     # noqa: E402

if __name__ == '__main__':
    parser = optparse.OptionParser(usage=__doc__)
    parser.add_option('--fs', type='float', default=69984000.0, help='sample rate of the recordings in Hz (default %default)')
    parser.add_option('--time', type='int', default=80, help='integration time in milliseconds (default %default)')
    (options, args) = parser.parse_args()
    if len(args) != 4:
        parser.error('need L1_FILE L2_FILE L5_FILE DEST_DIR')
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    sweep.run({1: args[0], 2: args[1], 3: args[2]}, options.fs, args[3], ms=options.time, rank=rank, world=world)
