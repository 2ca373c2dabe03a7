#!/usr/bin/env python
"""A/B of engine options on BASELINE config 2 (or cfg4 / cfg3 of tools/ab_v3.py): each option set is timed in
several interleaved rounds (CUDA events on the engine's stream, resident inputs, search only) and the records are
compared with the first set's. Usage: python tools/ab_quick.py [cfg2|cfg4|cfg3] "k=v,k=v" "k=v" ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]
import torch
from gnsstools import _native

CASES = {'cfg2': (163680, False, 32, 80, 1, True), 'cfg4': (30690, True, 64, 70, 20, False), 'cfg3': (81840, True, 72, 360, 1, False),
         'e6': (15345, True, 16, 30, 20, False)}
args = sys.argv[1:]
case = args.pop(0) if args and args[0] in CASES else 'cfg2'
sets = [dict((kv.split('=')[0], int(kv.split('=')[1])) for kv in a.split(',') if kv) for a in (args or [''])]
n, pad, R, D, B, norm = CASES[case]
dev = torch.device('cuda', 0)
eng = _native.Engine(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
N = 2 * n if pad else n
nx = (B - 1) * n + N
x = (rng.normal(0, 8, nx) + 1j * rng.normal(0, 8, nx)).astype(np.complex64)
rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1.0, -1.0).astype(np.float32)
if pad:
    rep[:, n:] = 0
eng.set_signal(x)
eng.set_replicas(rep)
f = -np.arange(-D // 2, D - D // 2) * 1e-5
keys = sorted(set(k for s in sets for k in s))
DEFAULTS = {'lanes': 2, 'fused_sets': 3, 'v3': 1, 'gt_split': 1, 'specialized_kernels': 1, 'overlap_chunks': 1}
rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
times = [[] for _ in sets]
ref = None
for rnd in range(4):
    for i, s in enumerate(sets):
        for k in keys:
            eng.set_option(k, s.get(k, DEFAULTS.get(k, 0)))
        eng.search_device(f, n, B, norm, 0, rec.data_ptr())
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if B == 1 else 4
        a.record(stream)
        for _ in range(reps):
            eng.search_device(f, n, B, norm, 0, rec.data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
        if rnd:
            times[i].append(a.elapsed_time(b) / reps)
        got = rec.cpu().numpy().copy()
        if ref is None:
            ref = got
        elif not np.array_equal(got.view(np.int32)[1::4], ref.view(np.int32)[1::4]):
            print('   !! lags differ from the first option set:', s)
for s, t in zip(sets, times):
    ms = float(np.median(t))
    print('%-6s %-60s %8.3f ms (min %.3f max %.3f)  %.3e cell-blocks/s' % (case, ' '.join('%s=%d' % kv for kv in sorted(s.items())) or '(defaults)',
                                                                      ms, min(t), max(t), R * D * N * B / ms * 1e3), flush=True)
eng.close()
