#!/usr/bin/env python
"""A few small searches covering every kernel family, for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool racecheck python tools/sanitize_cases.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]
from gnsstools import _native
eng = _native.Engine(0)
rng = np.random.default_rng(0)
# name, n, pad, R, D, B, normalize, spec
CASES = [('cube 4096', 4096, False, 2, 3, 2, True, 1), ('mid generic 4096', 4096, False, 1, 2, 2, True, 0),
         ('mid 2310 (3,5,7,11)', 2310, False, 1, 2, 1, False, 1), ('mid 4092 (31)', 4092, False, 1, 2, 1, False, 1),
         ('large 16384 spec', 8192, True, 2, 2, 2, False, 1), ('large 30690 spec (31,6 / 15,11)', 15345, True, 2, 2, 2, False, 1),
         ('large 61380 spec', 30690, True, 1, 2, 2, False, 1), ('large 163680 spec', 163680, False, 1, 2, 1, True, 1),
         ('large 163680 generic', 163680, False, 1, 1, 1, True, 0), ('large 50000 spec', 25000, True, 1, 2, 2, False, 1)]
def run(name, n, pad, R, D, B, norm, spec, opts=()):
    N = 2 * n if pad else n
    nx = (B - 1) * n + N
    x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1, -1).astype(np.int8)
    eng.set_option('specialized_kernels', spec)
    for k, v in opts:
        eng.set_option(k, v)
    eng.set_signal(x); eng.set_replicas(rep)
    m, l, d = eng.search(-np.arange(D) * 1e-5, n, B, norm)
    print(name, dict(opts), 'variant', eng.kernel_variant(), m[:2], l[:2], d[:2], flush=True)
    for k, _ in opts:
        eng.set_option(k, 3 if k == 'fused_sets' else 0)


for case in CASES:
    run(*case)
# round 2: coprime splits; copy-engine-fed pair in every tile shape (bulk copies, tensor-map copies, mbarriers,
# named barriers), the register-loading kernels on the same plans, the fused persistent kernel, embedded lengths
for rows, cols in ((0, 0), (1, 1), (2, 2), (3, 3), (4, 4), (5, 0)):      # rows 0: two-role kernel, 5: block-barrier 4-row kernel
    run('163680 copy-engine-fed pair', 163680, False, 3, 5, 1, True, 1, (('v3_rows', rows), ('v3_cols', cols), ('v3_rc', 2), ('v3_g', 2)))
run('61380 pair, 3 blocks', 30690, True, 2, 3, 3, False, 1, (('v3_cols', 3),))
run('61380 pair, 3 blocks', 30690, True, 2, 3, 3, False, 1)
run('30690 = 341 x 90, ragged column tiles', 15345, True, 2, 3, 2, False, 1)
run('163680 register-loading on the coprime split', 163680, False, 2, 2, 1, True, 1, (('v3', 0),))
eng.set_option('v3', 1)
run('163680 fused persistent kernel', 163680, False, 3, 5, 1, True, 1, (('fused', 1), ('fused_rc', 2), ('fused_g', 2), ('fused_sets', 2), ('fused_tpt', 7)))
run('61380 fused, 2 blocks', 30690, True, 2, 3, 2, False, 1, (('fused', 1),))
run('embedded 39406 -> 131072', 39406, False, 1, 2, 2, True, 1)
run('embedded 646 -> 2048', 646, False, 2, 2, 2, True, 1)
f = np.concatenate([-(np.arange(3) + 10 * g) * 1e-5 for g in range(4)])
print('grouped', eng.search_grouped(f, 3, 646, 2, True)[1].ravel()[:4])
print('sharded (no communicator)', eng.search_sharded(-np.arange(5) * 1e-5, 646, 2, True)[1])
raw = rng.integers(-127, 128, 2 * 40000).astype(np.int8)
eng.preprocess(raw, -0.01, 0.0, np.ones(161) / 161, 1.25, 20000)
xs = (rng.normal(0, 8, 5000) + 1j * rng.normal(0, 8, 5000)).astype(np.complex64)
eng.mix(xs, 0.123, 0.0)
print('front end + mix done')
# replica builder and correlator bank
eng.set_option('specialized_kernels', 1)
chips = rng.integers(0, 2, (2, 1023)).astype(np.int8)
eng.set_signal((rng.normal(0, 8, 3 * 4092) + 1j * rng.normal(0, 8, 3 * 4092)).astype(np.complex64))
eng.set_replicas_from_chips(chips, 4092, 8184, 1023.0 / 4092, boc=True)
print('builder', eng.search(-np.arange(2) * 1e-5, 4092, 2, False)[1])
long_code = rng.integers(0, 2, 767250).astype(np.int8)
base = rng.uniform(0, 767250, (5, 2))
print('bank', np.abs(eng.correlate_bank(long_code, -1e-3, 6000, 2, 6000, base, 0.125))[:, 0])
xe = (rng.integers(-60, 61, (2, 5000)) + 1j * rng.integers(-60, 61, (2, 5000))).astype(np.complex64)
ce = rng.integers(0, 2, (2, 10230)).astype(np.int8)
for mode, prm in ((0, None), (1, [1.0, -1.0, 0, 0]), (2, [1.0, -1.0, 0.953463, 0.301511]), (3, [1.0, -1.0, 0, 0] + [1, 0, 0, 0, 1, 0, 1] + [0] * 26)):
    print('epl mode', mode, np.abs(eng.correlate_epl(xe, ce, [100.2, 100.25, 100.3, -0.4], 0.2046, xsel=[0, 0, 1, 1], csel=[0, 1, 0, 1], mode=mode, params=prm))[:2])
