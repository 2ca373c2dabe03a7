"""Generate tests/golden/*.npz by running the REFERENCE's own search() (lifted from
/root/reference/acquire-*.py, see oracle/ref_lift.py) on seeded synthetic captures.
Run in the build container:  python tests/golden/make_golden.py
The fixtures are what the GPU box checks against, since /root/reference does not travel."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]

from oracle import ref_lift          # noqa: E402
from gnsstools import synth          # noqa: E402


def config1():
    """BASELINE config 1: acquire-gps-l1.py, PRN 1 planted, 1 ms, +-5 kHz / 500 Hz, all 32 PRNs."""
    ref_search, _ = ref_lift.lift_search('gps-l1')
    ms, seed, grid = 1, 1234, (-5000.0, 5000.0, 500.0)
    x = synth.capture('gps-l1', ms=ms, sats=[(1, 1500.0, 300.25, 4.0)], seed=seed)
    prns = np.arange(1, 33)
    out = [ref_search(x.astype(np.complex128), int(p), grid, ms) for p in prns]
    np.savez_compressed(os.path.join(HERE, 'gps_l1_config1.npz'), x=x, ms=ms, seed=seed, grid=np.array(grid),
                        prns=prns, metric=np.array([o[0] for o in out], np.float64),
                        code=np.array([o[1] for o in out], np.float64),
                        doppler=np.array([o[2] for o in out], np.float64))
    print('config1', out[0])


if __name__ == '__main__':
    config1()
