#!/usr/bin/env python
"""Throughput of the search on the other BASELINE / script-default shapes (resident inputs,
CUDA events on the engine's stream). Not the headline bench: a coverage check."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]
import torch
from gnsstools import _native

CASES = [
    # name, n, pad, R, D, B, normalize
    ('config1 gps-l1 PRN1 1ms', 4096, False, 1, 20, 1, True),
    ('gps-l1 defaults 32PRN 70D 80blk', 4096, False, 32, 70, 80, True),
    ('glonass-l1 1chan 70D 80blk', 16384, False, 1, 70, 80, False),
    ('glonass-l1 defaults 15chan x 70D 80blk, one grouped call', 16384, False, 1, 15 * 70, 80, False),
    ('b1i 63PRN 70D 80blk', 8192, True, 63, 70, 80, False),
    ('config3 e1b+e1c ref-style 65536', 32768, True, 72, 360, 1, False),
    ('e1b defaults 50PRN 360D 19blk', 32768, True, 50, 360, 19, False),
    ('l1cd defaults 32PRN 700D 8blk', 81920, False, 32, 700, 8, False),
    ('config4 l5i+l5q ref-style 61380', 30690, True, 64, 70, 20, False),
    ('l5i defaults 32PRN 70D 80blk', 30690, True, 32, 70, 80, False),
    ('config4 native 50000', 25000, True, 64, 70, 20, False),
    ('e6b defaults 50PRN 90D 80blk', 15345, True, 50, 90, 80, False),
    ('l2cm defaults 32PRN 700D 3blk', 81920, True, 32, 700, 3, False),
    ('config3 native 163680 padded', 81840, True, 72, 360, 1, False),
]

dev = torch.device('cuda', 0)
eng = _native.Engine(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
only = [a for a in sys.argv[1:] if '=' not in a]
for a in sys.argv[1:]:
    if '=' in a:                      # engine option name=value (A/B measurements)
        k, v = a.split('=')
        eng.set_option(k, int(v))
for name, n, pad, R, D, B, norm in CASES:
    if only and not any(o in name for o in only):
        continue
    N = 2 * n if pad else n
    nx = (B - 1) * n + N
    x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1.0, -1.0).astype(np.float32)
    if pad:
        rep[:, n:] = 0
    f = -np.arange(-D // 2, D - D // 2) * 1e-5
    eng.set_signal(x)
    eng.set_replicas(rep)
    rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
    def step():
        if 'grouped' in name:
            eng.search_grouped(f, 70, n, B, norm)            # the FDMA channel loop as Doppler groups (host results, synchronises)
        else:
            eng.search_device(f, n, B, norm, 0, rec.data_ptr())
    step(); torch.cuda.synchronize()
    reps = 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        step()
    b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    cells = R * D * N
    print('%-36s N=%-6d plan=%-22s variant=%d  %9.3f ms  %.3e cells/s  %.3e cell-blocks/s' % (
        name, N, '%dx%d%s' % (eng.plan_info()['N1'], eng.plan_info()['N2'], ' large' if eng.plan_info()['large'] else ' mid'),
        eng.kernel_variant(), ms, cells / ms * 1e3, cells * B / ms * 1e3), flush=True)
