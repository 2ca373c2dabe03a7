#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r03i_rows_v4.log 2>&1 <<'PY'
import os, sys, itertools
import numpy as np
ROOT = os.getcwd()
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')]
import torch
from gnsstools import _native
dev = torch.device('cuda', 0)
eng = _native.Engine(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
rng = np.random.default_rng(0)
for name, n, pad, R, D, B, norm in [('cfg2', 163680, False, 32, 80, 1, True), ('cfg3', 81840, True, 72, 360, 1, False)]:
    N = 2 * n if pad else n
    x = (rng.normal(0, 8, (B - 1) * n + N) + 1j * rng.normal(0, 8, (B - 1) * n + N)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, N)) > 0, 1.0, -1.0).astype(np.float32)
    if pad: rep[:, n:] = 0
    eng.set_signal(x); eng.set_replicas(rep)
    f = -np.arange(-D // 2, D - D // 2) * 1e-5
    rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
    ref = None
    for rows, cols, (rc, g), lanes in itertools.product((0, 4), (0, 1), ((0, 0), (16, 4), (32, 4), (32, 8)), (2, 3)):
        for k, v in (('v3_rows', rows), ('v3_cols', cols), ('v3_rc', rc), ('v3_g', g), ('lanes', lanes)):
            eng.set_option(k, v)
        eng.search_device(f, n, B, norm, 0, rec.data_ptr()); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5): eng.search_device(f, n, B, norm, 0, rec.data_ptr())
        b.record(stream); torch.cuda.synchronize()
        got = rec.cpu().numpy().copy()
        if ref is None: ref = got
        print('%s rows=%d cols=%d rc=%d g=%d lanes=%d  %.3f ms  %s' % (name, rows, cols, rc, g, lanes, a.elapsed_time(b) / 5, 'same' if np.array_equal(got, ref) else 'DIFFERENT'), flush=True)
PY
