"""Sweep driver (acquire-all equivalent): job sharding on the CPU, and on the GPU a reduced
sweep whose result lines equal what the single-signal path returns."""
import os

import numpy as np
import pytest

from gnsstools import sweep, acquire


def test_job_list_covers_reference_sweep():
    names = [j[1] for j in sweep.JOBS]
    assert len(sweep.JOBS) == 21 and len({j[3] for j in sweep.JOBS}) == 21
    assert all(n in acquire.SIGNALS for n in names)
    assert {j[0] for j in sweep.JOBS} == {1, 2, 3}


def test_sharding_is_a_partition_and_balanced():
    for world in (1, 2, 4, 8):
        parts = [sweep.shard_jobs(sweep.JOBS, 80, r, world) for r in range(world)]
        flat = [j for p in parts for j in p]
        assert sorted(flat) == sorted(sweep.JOBS)
        loads = [sum(sweep.job_cost(j[1], 80) for j in p) for p in parts]
        if world <= 4:
            assert max(loads) <= 1.6 * (sum(loads) / world)


@pytest.mark.gpu
def test_reduced_sweep_on_gpu(tmp_path):
    from gnsstools import _native, acquire_cli
    import synth_files
    eng = _native.Engine(0)
    fs = 20000000.0
    # one recording reused for the three "bands": GLONASS channel 1 planted (synth_files case)
    p = tmp_path / 'band.iq'
    p.write_bytes(synth_files.recording('glonass-l1'))
    jobs = [(1, 'glonass-l1', 250000, 'a.dat'), (1, 'gps-l1', 0, 'b.dat'), (2, 'beidou-b1i', 0, 'c.dat')]
    over = {'glonass-l1': ([-1, 0, 1], [-1000.0, 1000.0, 250.0]), 'gps-l1': ([1, 2], [-1000.0, 1000.0, 500.0]),
            'beidou-b1i': ([1, 2, 3], [-400.0, 400.0, 200.0])}
    out = sweep.run({1: str(p), 2: str(p)}, fs, str(tmp_path / 'out'), ms=2, jobs=jobs, engine=eng, overrides=over)
    assert sorted(os.listdir(tmp_path / 'out')) == ['a.dat', 'b.dat', 'c.dat']
    lines = open(tmp_path / 'out' / 'a.dat').read().splitlines()
    assert lines[2].startswith('chan  1 doppler   250.0')
    # identical to the single-signal path
    raw = np.frombuffer(p.read_bytes(), dtype=np.int8)
    acquire_cli.preprocess(acquire.SIGNALS['gps-l1'], raw, fs, 0.0, 7, engine=eng)
    want = acquire.acquire('gps-l1', None, [1, 2], [-1000.0, 1000.0, 500.0], 2, engine=eng)
    assert [r for _, r in out['b.dat']] == want
    eng.close()
