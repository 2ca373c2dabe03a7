#!/usr/bin/env python
"""A/B of four-step splits N = N1 x N2 (option split_n1) on the GPU: cell-blocks/s per split."""
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
CASES = [('163680 R32 D80 B1', 163680, False, 32, 80, 1, [372, 248, 496, 330, 440, 660]),
         ('61380 R32 D70 B20', 30690, True, 32, 70, 20, [220, 186, 279, 330]),
         ('30690 R50 D90 B20', 15345, True, 50, 90, 20, [165, 186]),
         ('65536 R72 D360 B1', 32768, True, 72, 360, 1, [256, 128, 512]),
         ('81920 R32 D100 B8', 81920, False, 32, 100, 8, [256, 320])]
for name, n, pad, R, D, B, splits in CASES:
    N = 2 * n if pad else n
    nx = (B - 1) * n + N
    x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1.0, -1.0).astype(np.float32)
    f = -np.arange(-D // 2, D - D // 2) * 1e-5
    rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
    for n1 in splits:
        eng.set_option('split_n1', n1)
        eng.set_signal(x); eng.set_replicas(rep)
        eng.search_device(f, n, B, False, 0, rec.data_ptr()); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(3):
            eng.search_device(f, n, B, False, 0, rec.data_ptr())
        b.record(stream); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        pi = eng.plan_info()
        print('%-20s n1=%-4d plan=%dx%d variant=%d %9.3f ms %.3e cell-blocks/s' % (name, n1, pi['N1'], pi['N2'], eng.kernel_variant(), ms, R * D * N * B / ms * 1e3), flush=True)
