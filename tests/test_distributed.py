"""Host logic of the multi-GPU path on CPU: world_size-2 gloo all-gather of per-rank records
(each rank runs the host-compiled kernels on its Doppler shard) + the strict-'>' merge must
reproduce the single-call answer bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnsstools import distributed as gd
from gnsstools._native import RECORD_DTYPE


def test_doppler_shard_partitions():
    for D in (1, 2, 7, 80, 81):
        for world in (1, 2, 3, 4, 8):
            cuts = [gd.doppler_shard(D, k, world) for k in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == D
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_merge_tie_goes_to_lowest_doppler():
    rec = np.zeros((3, 2), RECORD_DTYPE)
    rec['metric'] = [[5.0, 0.0], [5.0, 0.0], [7.0, 0.0]]
    rec['lag'] = [[10, 0], [20, 0], [30, 0]]
    rec['dbin'] = [[3, -1], [0, -1], [1, -1]]
    best = gd.merge_records(rec, [0, 4, 8])
    assert (best['metric'][0], best['lag'][0], best['dbin'][0]) == (7.0, 30, 9)
    assert best['dbin'][1] == -1                         # nothing > 0 anywhere -> reference returns (0,0,0)
    best2 = gd.merge_records(rec[:2], [0, 4])
    assert (best2['lag'][0], best2['dbin'][0]) == (10, 3)  # equal metric: first (lowest Doppler) wins


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import emu_util
    from oracle import acq_oracle as orc
    import gnsstools.gps.ca as ca
    n, fs = 1024, 1.024e6
    rng = np.random.default_rng(4)
    x = (rng.normal(0, 8, 3 * n) + 1j * rng.normal(0, 8, 3 * n))
    x += 3 * orc.resample_code(ca.ca_code(2), 100.0, 0, 1023.0 / n, 3 * n) * np.exp(2j * np.pi * 250.0 * np.arange(3 * n) / fs)
    x = x.astype(np.complex64)
    eng = emu_util.emu_engine()
    eng.set_signal(x)
    eng.set_replicas(np.array([orc.replica(ca.ca_code(p), n, False, False) for p in (1, 2, 3)]))
    f = -orc.doppler_bins((-1000, 1000, 250)) / fs
    m, l, d = gd.sharded_search(eng, f, n, 2, True, 0, rank, world, device=torch.device('cpu'))
    if rank == 0:
        m1, l1, d1 = eng.search(f, n, 2, True)
        ret['ok'] = bool(np.array_equal(m, m1) and np.array_equal(l, l1) and np.array_equal(d, d1) and d[1] == 5)
    eng.close()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_search():
    import emu_util
    emu_util.emu_cdll()                    # build once before forking
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get('ok') is True
