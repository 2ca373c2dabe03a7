#!/usr/bin/env python
"""A/B of planner options (disable_radix, split_n1, ...) on the GPU: cell-blocks/s per option set."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]
import torch
from gnsstools import _native
dev = torch.device('cuda', 0)
eng = _native.Engine(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
OPTS = [[], [('disable_radix', 22)], [('disable_radix', 20)], [('disable_radix', 20), ('disable_radix', 22)],
        [('disable_radix', 20), ('disable_radix', 22), ('disable_radix', 15)]]
CASES = [('163680 R32 D80 B1', 163680, False, 32, 80, 1),
         ('61380 R32 D70 B20', 30690, True, 32, 70, 20),
         ('30690 R50 D90 B20', 15345, True, 50, 90, 20),
         ('81920 R32 D100 B8', 81920, False, 32, 100, 8),
         ('163840 R32 D100 B3', 81920, True, 32, 100, 3),
         ('50000 R64 D70 B20', 25000, True, 64, 70, 20)]
for name, n, pad, R, D, B in CASES:
    N = 2 * n if pad else n
    nx = (B - 1) * n + N
    x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1.0, -1.0).astype(np.float32)
    f = -np.arange(-D // 2, D - D // 2) * 1e-5
    rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
    for opts in OPTS:
        eng.set_option('disable_radix', 0)
        for k, v in opts:
            eng.set_option(k, v)
        eng.set_signal(x); eng.set_replicas(rep)
        eng.search_device(f, n, B, False, 0, rec.data_ptr()); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(3):
            eng.search_device(f, n, B, False, 0, rec.data_ptr())
        b.record(stream); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        pi = eng.plan_info()
        print('%-20s %-34s plan=%dx%d variant=%d %9.3f ms %.3e cell-blocks/s' % (name, ','.join('-%d' % v for _, v in opts) or 'default', pi['N1'], pi['N2'], eng.kernel_variant(), ms, R * D * N * B / ms * 1e3), flush=True)
