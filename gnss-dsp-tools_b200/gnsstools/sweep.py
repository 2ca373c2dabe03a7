"""Multi-constellation acquisition sweep in one process (BASELINE config 5).

The reference's acquire-all.sh (acquire-all.sh:9-35) starts one Python interpreter per signal
and re-reads the three-band recording through the external packet2wav_3ch for each of them.
Here the band recordings are read once, every signal's front end and search run on the GPU
against them, and the per-signal result files (acq-<signal>.dat, same lines as the scripts
print) are written at the end. Under torchrun the job list is sharded across ranks by
estimated cost (signals differ 100x); ranks are independent, no collective is needed.
"""

import math
import os

import numpy as np

from . import acquire as acq
from . import acquire_cli, util, _native

# (band, signal, carrier offset Hz, output name) — acquire-all.sh:9-35, recording centre
# frequencies L1 1584.754875 MHz, L2 1227.727125 MHz, L5 1191.641625 MHz at 69.984 Msps.
# The shell script's glonass-l3i/l3q entries name scripts that do not exist in the reference;
# the L3OC data/pilot scripts that do exist are run in their place.
JOBS = [
    (1, 'gps-l1', -9334875, 'acq-gps-l1.dat'),
    (1, 'glonass-l1', 17245125, 'acq-glonass-l1.dat'),
    (1, 'galileo-e1b', -9334875, 'acq-galileo-e1b.dat'),
    (1, 'galileo-e1c', -9334875, 'acq-galileo-e1c.dat'),
    (1, 'beidou-b1i', -23656875, 'acq-beidou-b1i.dat'),
    (2, 'gps-l2cm', -127126, 'acq-gps-l2cm.dat'),
    (2, 'glonass-l2', 18272874, 'acq-glonass-l2.dat'),
    (2, 'glonass-l3ocd', -25702126, 'acq-glonass-l3ocd.dat'),
    (2, 'glonass-l3ocp', -25702126, 'acq-glonass-l3ocp.dat'),
    (2, 'galileo-e5bi', -20587126, 'acq-galileo-e5bi.dat'),
    (2, 'galileo-e5bq', -20587126, 'acq-galileo-e5bq.dat'),
    (2, 'beidou-b2i', -20587126, 'acq-beidou-b2i.dat'),
    (3, 'gps-l5i', -15191625, 'acq-gps-l5i.dat'),
    (3, 'gps-l5q', -15191625, 'acq-gps-l5q.dat'),
    (3, 'galileo-e5ai', -15191625, 'acq-galileo-e5ai.dat'),
    (3, 'galileo-e5aq', -15191625, 'acq-galileo-e5aq.dat'),
    (3, 'glonass-l3ocd', 10383375, 'acq-glonass-l3ocd-ch3.dat'),
    (3, 'glonass-l3ocp', 10383375, 'acq-glonass-l3ocp-ch3.dat'),
    (3, 'galileo-e5bi', 15498375, 'acq-galileo-e5bi-ch3.dat'),
    (3, 'galileo-e5bq', 15498375, 'acq-galileo-e5bq-ch3.dat'),
    (3, 'beidou-b2i', 15498375, 'acq-beidou-b2i-ch3.dat'),
]


def default_keys(sig):
    if sig.fdma:
        return util.parse_list_ranges(sig.prns, sep=':')
    return util.parse_list_ranges(sig.prns)


def job_cost(signal, ms):
    """Relative cost of one job: R * D * B * N * log2 N cell-block-stages."""
    sig = acq.SIGNALS[signal]
    R = len(default_keys(sig))
    D = len(acq.doppler_bins(util.parse_list_floats(sig.doppler)))
    B = max(sig.blocks(ms), 0)
    return R * D * B * sig.N * math.log2(sig.N)


def shard_jobs(jobs, ms, rank, world):
    """Greedy longest-first assignment of jobs to ranks; returns this rank's jobs in list order."""
    order = sorted(range(len(jobs)), key=lambda i: -job_cost(jobs[i][1], ms))
    load = [0.0] * world
    owner = {}
    for i in order:
        k = min(range(world), key=lambda r: load[r])
        owner[i] = k
        load[k] += job_cost(jobs[i][1], ms)
    return [jobs[i] for i in range(len(jobs)) if owner[i] == rank]


def run(band_files, fs, dest_dir, ms=80, jobs=None, engine=None, rank=0, world=1, overrides=None):
    """band_files: {band number: path of int8 I/Q recording}. Writes dest_dir/<out name> per job
    and returns {out name: [(key, (metric, code, doppler)), ...]}. `overrides` maps a signal name
    to (keys, doppler_search) for reduced sweeps."""
    eng = engine if engine is not None else _native.default_engine()
    jobs = shard_jobs(JOBS if jobs is None else jobs, ms, rank, world)
    os.makedirs(dest_dir, exist_ok=True)
    ms_pad = ms + 5
    n = int(fs * 0.001 * ms_pad)
    raw = {}
    out = {}
    for band, signal, coffset, name in jobs:
        if band not in band_files:
            continue
        if band not in raw:
            with open(band_files[band], 'rb') as fp:
                raw[band] = np.frombuffer(fp.read(2 * n), dtype=np.int8)
            if raw[band].size != 2 * n:
                raise TypeError('short read: %s' % band_files[band])
        sig = acq.SIGNALS[signal]
        keys, grid = (overrides or {}).get(signal, (default_keys(sig), util.parse_list_floats(sig.doppler)))
        acquire_cli.preprocess(sig, raw[band], fs, float(coffset), ms_pad, engine=eng)
        res = acq.acquire(signal, None, keys, grid, ms, engine=eng)
        with open(os.path.join(dest_dir, name), 'w') as f:
            for key, r in zip(keys, res):
                f.write(acq.format_result(signal, key, r) + '\n')
        out[name] = list(zip(keys, res))
    return out
