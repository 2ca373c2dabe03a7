"""Multi-GPU through the C ABI alone (no torch.distributed): gnssacq_nccl_unique_id /
gnssacq_nccl_init / gnssacq_search_sharded. One process per GPU, the ncclUniqueId travels through a
file; every rank must return the single-GPU answer bit for bit (contiguous
ascending Doppler shards, one all-gather, rank-ordered strict-'>' merge).
Also runnable by hand on a multi-GPU box:  python tests/test_nccl_abi.py 4"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, 'gnss-dsp-tools_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)


def _inputs():
    rng = np.random.default_rng(42)
    n, R, D, B = 16384, 5, 23, 2
    x = (rng.normal(0, 8, (B + 1) * n) + 1j * rng.normal(0, 8, (B + 1) * n)).astype(np.complex64)
    rep = np.where(rng.integers(0, 2, (R, n)) > 0, 1, -1).astype(np.int8)
    x[:2 * n] += 2 * rep[3].astype(np.float32).repeat(1)[np.arange(2 * n) % n] * np.exp(2j * np.pi * 7e-5 * np.arange(2 * n))
    f = -(np.arange(D) - D // 2) * 1e-5
    return x, rep, f, n, B


def _worker(rank, world, idfile, outdir):
    from gnsstools import _native
    eng = _native.Engine(rank)
    if rank == 0:
        uid = eng.nccl_unique_id()
        with open(idfile + '.tmp', 'wb') as fp:
            fp.write(uid)
        os.rename(idfile + '.tmp', idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise RuntimeError('no ncclUniqueId from rank 0')
            time.sleep(0.05)
        uid = open(idfile, 'rb').read()
    eng.nccl_init(uid, rank, world)
    x, rep, f, n, B = _inputs()
    eng.set_signal(x)
    eng.set_replicas(rep)
    for normalize in (False, True):
        got = eng.search_sharded(f, n, B, normalize)
        np.savez(os.path.join(outdir, 'rank%d_%d.npz' % (rank, int(normalize))), m=got[0], l=got[1], d=got[2])
    if rank == 0:
        for normalize in (False, True):
            m, l, d = eng.search(f, n, B, normalize)
            np.savez(os.path.join(outdir, 'single_%d.npz' % int(normalize)), m=m, l=l, d=d)
    eng.close()


def run_world(world, outdir):
    ctx = mp.get_context('spawn')
    idfile = os.path.join(outdir, 'nccl_id')
    procs = [ctx.Process(target=_worker, args=(r, world, idfile, outdir)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for normalize in (0, 1):
        want = np.load(os.path.join(outdir, 'single_%d.npz' % normalize))
        assert int(want['d'][3]) == 23 // 2 + 7 and want['m'][3] > 2 * np.median(want['m'])      # the planted replica, at its bin
        for r in range(world):
            got = np.load(os.path.join(outdir, 'rank%d_%d.npz' % (r, normalize)))
            for k in ('m', 'l', 'd'):
                assert np.array_equal(got[k], want[k]), (world, r, normalize, k, got[k], want[k])


@pytest.mark.gpu
@pytest.mark.parametrize('world', [1, 2, 4, 8])
def test_sharded_search_through_the_c_abi(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    run_world(world, str(tmp_path))


def test_sharded_search_without_a_communicator_is_the_plain_search():
    """CPU: the emulated kernels through gnssacq_search_sharded with no communicator (world 1)."""
    import emu_util
    eng = emu_util.emu_engine()
    x, rep, f, n, B = _inputs()
    eng.set_signal(x[:3 * 2048])
    eng.set_replicas(rep[:, :2048])
    a = eng.search(f, 2048, 2, True)
    b = eng.search_sharded(f, 2048, 2, True)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    eng.close()


if __name__ == '__main__':
    import tempfile
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    with tempfile.TemporaryDirectory() as d:
        run_world(w, d)
    print('sharded search through the C ABI on %d GPUs: every rank returned the single-GPU answer' % w)
