"""Build the data files the code generators need from the ICD tables carried by the
reference checkout, and the golden chip hashes the tests check the generators against.

Run in the build container (needs /root/reference):  python tools/extract_code_tables.py

Outputs
  gnss-dsp-tools_b200/gnsstools/_data/icd_tables.json   per-PRN generator parameters (ICD tables)
  gnss-dsp-tools_b200/gnsstools/_data/memory_codes.npz  memory codes / tabulated secondary codes, bit-packed
  tests/golden/code_hashes.json                        sha256 of every PRN's chips as the REFERENCE generates them
Only data (interface-specification constants) is extracted; the generator algorithms in
gnsstools/* are written independently.
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_lift  # noqa: E402

warnings.filterwarnings('ignore')
DATA = os.path.join(ROOT, 'gnss-dsp-tools_b200', 'gnsstools', '_data')

TABLES = {
    'gps.l1cd': ['l1cd_params'], 'gps.l1cp': ['l1cp_params', 'l1cp_secondary_params'],
    'gps.l2cm': ['l2cm_init', 'l2cm_end_state'], 'gps.l2cl': ['l2cl_init', 'l2cl_end_state'],
    'gps.l5i': ['l5i_init'], 'gps.l5q': ['l5q_init'],
    'galileo.e5ai': ['e5ai_init'], 'galileo.e5aq': ['e5aq_init'], 'galileo.e5bi': ['e5bi_init'],
    'galileo.e5bq': ['e5bq_init'],
    'beidou.b1i': ['b1i_g2_taps'], 'beidou.b1cd': ['b1cd_params'],
    'beidou.b1cp': ['b1cp_params', 'b1cp_secondary_params'],
    'beidou.b2ad': ['b2ad_g2_initial'], 'beidou.b2ap': ['b2ap_g2_initial', 'b2ap_secondary_params'],
    'beidou.b2bd': ['b2bd_g2_initial'], 'beidou.b2bp': ['b2bp_g2_initial'], 'beidou.b3i': ['b3i_g2_initial'],
}
MEMORY = ['galileo.e1b', 'galileo.e1c', 'galileo.e6b', 'galileo.e6c', 'beidou.b2bi', 'beidou.b2bq',
          'xona.x1d', 'xona.x1p', 'xona.x5p']
SECONDARY_TABLES = ['galileo.e5aq', 'galileo.e5bq', 'galileo.e6c']
ALL = ['gps.ca', 'gps.l1cd', 'gps.l1cp', 'gps.l2cm', 'gps.l2cl', 'gps.l5i', 'gps.l5q',
       'galileo.e1b', 'galileo.e1c', 'galileo.e5ai', 'galileo.e5aq', 'galileo.e5bi', 'galileo.e5bq',
       'galileo.e6b', 'galileo.e6c',
       'beidou.b1i', 'beidou.b1cd', 'beidou.b1cp', 'beidou.b2ad', 'beidou.b2ap', 'beidou.b2bd', 'beidou.b2bp',
       'beidou.b2bi', 'beidou.b2bq', 'beidou.b3i',
       'glonass.ca', 'glonass.p', 'glonass.l3ocd', 'glonass.l3ocp', 'xona.x1d', 'xona.x1p', 'xona.x5p']


def ref(mod):
    return ref_lift.ref_import('gnsstools.' + mod)


def code_fn(mod):
    return getattr(ref(mod), mod.split('.')[-1] + '_code')


def prn_list(mod):
    m = ref(mod)
    name = mod.split('.')[-1]
    for key in (name + '_params', name + '_init', name + '_g2_taps', name + '_g2_initial', name + '_strings', 'g2_delay'):
        if hasattr(m, key):
            return sorted(getattr(m, key).keys())
    if mod in ('glonass.l3ocd', 'glonass.l3ocp'):
        return list(range(64))
    return None     # single code, no PRN argument


def digest(a):
    return hashlib.sha256(np.asarray(a).astype(np.uint8).tobytes()).hexdigest()[:24]


def main():
    os.makedirs(DATA, exist_ok=True)
    tables = {}
    for mod, names in TABLES.items():
        m = ref(mod)
        tables[mod] = {n: {str(k): (list(v) if isinstance(v, tuple) else v) for k, v in getattr(m, n).items()} for n in names}
    for mod in ('xona.x1p', 'xona.x5p'):          # fixed 100-chip secondary codes
        sc = ref(mod).secondary_code
        tables[mod] = {'secondary_bits': {'0': ''.join('%d' % int((1.0 - v) / 2.0) for v in sc)}}
    with open(os.path.join(DATA, 'icd_tables.json'), 'w') as f:
        json.dump(tables, f, separators=(',', ':'), sort_keys=True)

    packed = {}
    for mod in MEMORY:
        prns = prn_list(mod)
        fn = code_fn(mod)
        bits = np.array([fn(p) for p in prns]).astype(np.uint8)
        packed[mod + ':prns'] = np.array(prns, np.int32)
        packed[mod + ':length'] = np.array(bits.shape[1], np.int32)
        packed[mod + ':bits'] = np.packbits(bits, axis=1)
    for mod in SECONDARY_TABLES:
        sc = ref(mod).secondary_code           # dict prn -> +-1 float array (100 chips)
        prns = sorted(sc.keys())
        bits = np.array([(1.0 - sc[p]) / 2.0 for p in prns]).astype(np.uint8)
        packed[mod + ':sec_prns'] = np.array(prns, np.int32)
        packed[mod + ':sec_length'] = np.array(bits.shape[1], np.int32)
        packed[mod + ':sec_bits'] = np.packbits(bits, axis=1)
    np.savez_compressed(os.path.join(DATA, 'memory_codes.npz'), **packed)

    hashes = {}
    for mod in ALL:
        prns = prn_list(mod)
        fn = code_fn(mod)
        h = {}
        if prns is None:
            h['-'] = digest(fn())
        else:
            if mod == 'gps.l2cl':
                prns = [1, 2, 32, 159]          # 767250-chip Python loops: ~20 s per PRN in the reference
            for p in prns:
                h[str(p)] = digest(fn(p))
        m = ref(mod)
        sec = {}
        if hasattr(m, 'secondary_code') and not callable(m.secondary_code):
            sc = m.secondary_code
            if isinstance(sc, dict):
                sec = {str(p): digest((1.0 - sc[p]) / 2.0) for p in sorted(sc)}
            else:
                sec = {'-': digest((1.0 - np.asarray(sc)) / 2.0)}
        elif hasattr(m, 'secondary_code'):
            for p in (prn_list(mod) or []):
                sec[str(p)] = digest(m.secondary_code(p))
        hashes[mod] = {'code_length': int(m.code_length), 'chip_rate': int(m.chip_rate), 'codes': h, 'secondary': sec}
        print(mod, len(h), 'codes', len(sec), 'secondary', flush=True)
    with open(os.path.join(ROOT, 'tests', 'golden', 'code_hashes.json'), 'w') as f:
        json.dump(hashes, f, indent=0, sort_keys=True)


if __name__ == '__main__':
    main()
