#!/usr/bin/env python
"""A/B of explicit stage schedules (gnssacq_set_schedule) on the GPU."""
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
CASES = [('163680 R32 D80 B1', 163680, False, 32, 80, 1,
          [([31, 12], [20, 22]), ([31, 12], [11, 5, 8]), ([31, 3, 4], [11, 5, 8]), ([31, 4, 3], [11, 5, 8]), ([31, 12], [5, 8, 11]),
           ([31, 12], [8, 5, 11]), ([31, 12], [11, 8, 5]), ([31, 3, 4], [11, 8, 5])]),
         ('81920 R32 D100 B8', 81920, False, 32, 100, 8, [([16, 16], [5, 8, 8]), ([16, 16], [8, 8, 5]), ([16, 16], [16, 20])]),
         ('163840 R32 D100 B3', 81920, True, 32, 100, 3, [([5, 8, 8], [8, 8, 8]), ([8, 8, 5], [8, 8, 8]), ([16, 20], [8, 8, 8])])]
for name, n, pad, R, D, B, scheds in CASES:
    N = 2 * n if pad else n
    nx = (B - 1) * n + N
    x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1.0, -1.0).astype(np.float32)
    f = -np.arange(-D // 2, D - D // 2) * 1e-5
    rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
    for s1, s2 in scheds:
        eng.set_schedule(1, s1); eng.set_schedule(2, s2)
        eng.set_signal(x); eng.set_replicas(rep)
        eng.search_device(f, n, B, False, 0, rec.data_ptr()); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5):
            eng.search_device(f, n, B, False, 0, rec.data_ptr())
        b.record(stream); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        pi = eng.plan_info()
        print('%-20s N1:%-12s N2:%-12s variant=%d %9.3f ms %.3e cell-blocks/s' % (name, s1, s2, eng.kernel_variant(), ms, R * D * N * B / ms * 1e3), flush=True)
