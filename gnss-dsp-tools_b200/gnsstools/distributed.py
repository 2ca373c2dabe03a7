"""Multi-GPU sharding of the acquisition search: one process per GPU (torchrun), the
Doppler grid split into contiguous ascending ranges, one all-gather of the per-replica
records, and the same strict-'>' reduction on every rank.

The reference's only parallelism is mp.Pool over PRNs (acquire-gps-l1.py:105-108); here
every (replica, Doppler) cell is independent until the final per-replica best, so the only
exchange is R 16-byte records per rank — latency-bound, plain NCCL all-gather.
"""

import numpy as np

from ._native import RECORD_DTYPE


def doppler_shard(D, rank, world):
    """Contiguous ascending slice of D Doppler bins owned by `rank` (sizes differ by <= 1)."""
    base, extra = divmod(D, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_records(all_rec, offsets):
    """Fold per-rank records (world x R structured array, rank order = ascending Doppler)
    into the single-GPU answer: a later rank replaces the best only on a strictly greater
    metric, so ties go to the lowest Doppler bin exactly as the reference's ascending scan
    (acquire-gps-l1.py:26,36). offsets[k] is the global index of rank k's first bin."""
    all_rec = np.asarray(all_rec)
    best = all_rec[0].copy()
    best['dbin'] = np.where(best['dbin'] >= 0, best['dbin'] + offsets[0], -1)
    for k in range(1, all_rec.shape[0]):
        rec = all_rec[k].copy()
        rec['dbin'] = np.where(rec['dbin'] >= 0, rec['dbin'] + offsets[k], -1)
        take = (rec['dbin'] >= 0) & (rec['metric'] > best['metric'])
        best = np.where(take, rec, best)
    return best


def allgather_records(rec, world, group=None):
    """rec: int32 torch tensor of 4*R words (R gnssacq_record_t) on this rank's device.
    Returns a (world, R) structured numpy array, identical on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        out = rec
    else:
        out = torch.empty(world * rec.numel(), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, rec, group=group)
    host = out.cpu().numpy()
    return host.view(RECORD_DTYPE).reshape(world, -1)


def sharded_search(engine, nco_freq, block_stride, n_blocks, normalize, n_lags, rank, world, device=None):
    """Search this rank's Doppler shard, all-gather, merge. Returns (metric, lag, dbin) for
    all R replicas with dbin indexing the full nco_freq list. GPU path: records stay on the
    device until after the collective."""
    import torch
    lo, hi = doppler_shard(len(nco_freq), rank, world)
    R = engine.R
    dev = device if device is not None else torch.device('cuda', torch.cuda.current_device())
    rec = torch.zeros(4 * R, dtype=torch.int32, device=dev)
    if hi > lo:
        engine.search_device(np.ascontiguousarray(nco_freq[lo:hi]), block_stride, n_blocks, normalize, n_lags, rec.data_ptr())
        engine.synchronize()
    else:
        rec.view(-1, 4)[:, 2] = -1            # empty shard: nothing selected
    allrec = allgather_records(rec, world)
    offsets = [doppler_shard(len(nco_freq), k, world)[0] for k in range(world)]
    best = merge_records(allrec, offsets)
    return best['metric'].copy(), best['lag'].copy(), best['dbin'].copy()
