"""The host-side module surface the reference scripts import — gnsstools.nco.nco / boc11,
gnsstools.io.get_samples_complex, gnsstools.util — against vectors produced by the reference's
own modules (tests/golden/host_surface.npz, made by tests/golden/make_host_golden.py), and
directly against the reference when /root/reference is present."""
import io
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from make_host_golden import NCO_ARGS, BOC_ARGS, RANGES, CHANNELS, FLOATS   # noqa: E402  (argument sets only)

from gnsstools import nco, util
from gnsstools import io as gio

G = np.load(os.path.join(HERE, 'golden', 'host_surface.npz'))
META = json.loads(str(G['meta']))


def same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize('i', range(len(NCO_ARGS)))
def test_nco_matches_reference_vectors(i):
    got = nco.nco(*NCO_ARGS[i])
    assert got.dtype == np.complex128 and same(got, G['nco_%d' % i])          # gnsstools/nco.py:6-10, bit for bit


@pytest.mark.parametrize('i', range(len(BOC_ARGS)))
def test_boc11_matches_reference_vectors(i):
    got = nco.boc11(*BOC_ARGS[i])
    assert same(got, G['boc_%d' % i]) and set(np.unique(got)) <= {-1, 1}       # gnsstools/nco.py:12-19


def test_get_samples_complex_matches_reference_vectors():
    raw = G['io_raw'].tobytes()
    got = gio.get_samples_complex(io.BytesIO(raw), 1001)
    assert got.dtype == np.complex64 and same(got, G['io_1001'])               # gnsstools/io.py:3-12
    fp = io.BytesIO(raw)                                                         # consecutive reads continue in the file
    assert same(gio.get_samples_complex(fp, 400), G['io_first_400'])
    assert same(gio.get_samples_complex(fp, 601), G['io_next_601'])
    assert gio.get_samples_complex(fp, 1) is None                                # exhausted
    assert gio.get_samples_complex(io.BytesIO(raw), 1002) is None                # short read -> None
    assert gio.get_samples_complex(io.BytesIO(b''), 0).size == 0                 # empty request, empty result
    # value mapping: I = even bytes, Q = odd bytes, signed
    z = gio.get_samples_complex(io.BytesIO(bytes([1, 255, 128, 127])), 2)
    assert list(z) == [1 - 1j, -128 + 127j]


def test_get_samples_complex_reads_a_real_file(tmp_path):
    raw = np.random.default_rng(5).integers(-128, 128, 2 * 4096, dtype=np.int8)
    p = tmp_path / 'rec.iq'
    p.write_bytes(raw.tobytes())
    with open(p, 'rb') as fp:
        x = gio.get_samples_complex(fp, 4096)
    assert np.array_equal(x.real, raw[0::2].astype(np.float32)) and np.array_equal(x.imag, raw[1::2].astype(np.float32))


def test_util_parsers_match_reference():
    for s, want in META['ranges'].items():
        assert util.parse_list_ranges(s) == want
    for s, want in META['channels'].items():
        assert util.parse_list_ranges(s, sep=':') == want
    for s, want in META['floats'].items():
        assert util.parse_list_floats(s) == want
    assert RANGES and CHANNELS and FLOATS


def test_host_surface_against_the_reference_modules_directly():
    from oracle import ref_lift
    if not ref_lift.available():
        pytest.skip('/root/reference not present')
    rnco = ref_lift.ref_import('gnsstools.nco')
    rio = ref_lift.ref_import('gnsstools.io')
    rng = np.random.default_rng(9)
    for _ in range(20):
        f, p, n = float(rng.uniform(-0.5, 0.5)), float(rng.uniform(0, 1)), int(rng.integers(1, 5000))
        assert same(nco.nco(f, p, n), rnco.nco(f, p, n))
        c, fr, inc = float(rng.integers(0, 5000)), float(rng.uniform(0, 1)), float(rng.uniform(0.01, 1.2))
        assert same(nco.boc11(c, fr, inc, n), rnco.boc11(c, fr, inc, n))
        raw = rng.integers(-128, 128, 2 * n, dtype=np.int8).tobytes()
        assert same(gio.get_samples_complex(io.BytesIO(raw), n), rio.get_samples_complex(io.BytesIO(raw), n))
    assert np.array_equal(nco.nco_table, rnco.nco_table)
