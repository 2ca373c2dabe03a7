#!/usr/bin/env python
"""A/B of the copy-engine-fed correlate pair (kernels_v3.cuh) on the GPU: tile shapes, chunk shapes
and lanes, against the register-loading kernels (v3=0), on BASELINE config 2 and the 61380 shape.
Resident inputs, CUDA events on the engine's stream, search only (replica set-up excluded)."""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]
import torch
from gnsstools import _native

dev = torch.device('cuda', 0)
eng = _native.Engine(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
quick = 'quick' in sys.argv[1:]
nofused = 'nofused' in sys.argv[1:]
args = [a for a in sys.argv[1:] if a not in ('quick', 'nofused')]
which = args[0] if args else 'all'

CASES = [('cfg2 163680 R32 D80 B1', 163680, False, 32, 80, 1, True),
         ('cfg4 61380 R64 D70 B20', 30690, True, 64, 70, 20, False),
         ('cfg3 163680pad R72 D360 B1', 81840, True, 72, 360, 1, False)]


def run(name, n, pad, R, D, B, norm, opts, reps=5):
    for k, v in opts.items():
        eng.set_option(k, v)
    rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
    f = -np.arange(-D // 2, D - D // 2) * 1e-5
    eng.search_device(f, n, B, norm, 0, rec.data_ptr())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        eng.search_device(f, n, B, norm, 0, rec.data_ptr())
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    N = 2 * n if pad else n
    print('%-28s %-58s var=%3d %8.3f ms  %.3e cell-blocks/s' % (name, ' '.join('%s=%d' % kv for kv in opts.items()), eng.kernel_variant(), ms,
                                                             R * D * N * B / ms * 1e3), flush=True)
    return rec.cpu().numpy().copy()


for name, n, pad, R, D, B, norm in CASES:
    if which != 'all' and which not in name:
        continue
    N = 2 * n if pad else n
    nx = (B - 1) * n + N
    x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1.0, -1.0).astype(np.float32)
    if pad:
        rep[:, n:] = 0
    eng.set_signal(x)
    eng.set_replicas(rep)
    base = dict(v3=0, fused=0, v3_rows=0, v3_cols=0, v3_rc=0, v3_g=0, lanes=2)
    ref = run(name, n, pad, R, D, B, norm, base)
    nrows = 5 if N == 163680 else 2
    ncols = 5 if N == 163680 else 2
    if quick:
        nrows, ncols = 1, 2
    shapes = [(0, 0), (16, 4), (32, 4), (32, 8), (16, 16), (8, 16)] if B == 1 else [(0, 0), (16, 1), (32, 1), (64, 1)]
    # tile shapes at the default chunk shape
    for rv, cv in ([] if quick else itertools.product(range(nrows), range(ncols))):
        got = run(name, n, pad, R, D, B, norm, dict(base, v3=1, v3_rows=rv, v3_cols=cv))
        if not np.array_equal(got.view(np.int32)[1::4], ref.view(np.int32)[1::4]):
            print('   !! lags differ from the register-loading kernels')
    # the fused persistent kernel: group shapes, column tiles per ticket, ring depth
    fshapes = [(4, 4), (2, 8), (4, 8), (8, 4), (16, 2)] if B == 1 else [(4, 1), (8, 1), (16, 1)]
    for (rc, g), tpt, sets in ([] if nofused else itertools.product(fshapes, (1, 3, 5, 10, 15), (3,))):
        got = run(name, n, pad, R, D, B, norm, dict(v3=1, fused=1, fused_rc=rc, fused_g=g, fused_sets=sets, fused_tpt=tpt))
        if not np.array_equal(got.view(np.int32)[1::4], ref.view(np.int32)[1::4]):
            print('   !! lags differ from the register-loading kernels')
    if not nofused:
        run(name, n, pad, R, D, B, norm, dict(v3=1, fused=1, fused_rc=4, fused_g=8, fused_sets=4, fused_tpt=5))
    run(name, n, pad, R, D, B, norm, dict(v3=1, fused=0, fused_rc=0, fused_g=0, fused_sets=3, fused_tpt=0))
    if quick:
        continue
    # chunk shapes and lanes at the default tiles
    for (rc, g), lanes in itertools.product(shapes, (2, 3)):
        run(name, n, pad, R, D, B, norm, dict(base, v3=1, v3_rc=rc, v3_g=g, lanes=lanes))
eng.close()
