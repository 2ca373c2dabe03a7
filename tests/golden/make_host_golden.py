"""Golden vectors for the host-side module surface (gnsstools.nco.nco / boc11, gnsstools.io,
gnsstools.util), produced by importing the REFERENCE's modules from /root/reference.
Build container only:  python tests/golden/make_host_golden.py  -> tests/golden/host_surface.npz"""
import io as _io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]

from oracle import ref_lift          # noqa: E402

# the argument sets are shared with tests/test_host_surface.py
NCO_ARGS = [(-1500.0 / 4096000.0, 0, 4096), (0.01234, 0.25, 1000), (-0.49, 0.9, 777), (5000.0 / 16368000.0, 0, 16368),
            (-(562500.0 * -7 + 250.0) / 16384000.0, 0, 2048), (0.0, 0.0, 5), (1.0 / 1024, 0.999, 2049)]
BOC_ARGS = [(0, 0, 4092 / 32768., 32768), (1, 0.25, 10230 / 81920., 5000), (3.0, 0.5, 0.731, 999), (1023, 0, 1.0, 64)]
RANGES = ['1-32', '1,3,7-14,31', '0', '19-21,40']
CHANNELS = ['-7:7', '-6,-4,-1:2,7']
FLOATS = ['-7000,7000,200', '-5000,5000,500', '0.5,1e3,-2']


def main():
    rnco = ref_lift.ref_import('gnsstools.nco')
    rio = ref_lift.ref_import('gnsstools.io')
    rutil = ref_lift.ref_import('gnsstools.util')
    out = {}
    for i, a in enumerate(NCO_ARGS):
        out['nco_%d' % i] = rnco.nco(*a)
    for i, a in enumerate(BOC_ARGS):
        out['boc_%d' % i] = rnco.boc11(*a)
    raw = np.random.default_rng(77).integers(-128, 128, 2 * 1001, dtype=np.int8).tobytes()
    out['io_raw'] = np.frombuffer(raw, dtype=np.int8)
    out['io_1001'] = rio.get_samples_complex(_io.BytesIO(raw), 1001)
    fp = _io.BytesIO(raw)
    out['io_first_400'] = rio.get_samples_complex(fp, 400)
    out['io_next_601'] = rio.get_samples_complex(fp, 601)
    assert rio.get_samples_complex(_io.BytesIO(raw), 1002) is None          # short read -> None (gnsstools/io.py:5-6)
    meta = {'ranges': {s: rutil.parse_list_ranges(s) for s in RANGES},
            'channels': {s: rutil.parse_list_ranges(s, sep=':') for s in CHANNELS},
            'floats': {s: rutil.parse_list_floats(s) for s in FLOATS},
            'dtypes': {k: str(v.dtype) for k, v in out.items()}}
    out['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, 'host_surface.npz'), **out)
    print({k: (v.dtype, v.shape) for k, v in out.items()})


if __name__ == '__main__':
    main()
