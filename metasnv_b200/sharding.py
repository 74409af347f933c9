"""Multi-GPU plumbing: which shard runs where, and how per-rank timings are combined.

Shards of this path are independent (createOptimumSplit.py:43-60 bins whole genomes and metaSNV.py
runs one pipe per bin), so there is no data-path collective: torch.distributed is only used for the
barrier and for max(time) / sum(work) over ranks.
"""
import re


def device_for_split(split_name, n_devices):
    """GPU of a `best_split_<k>` shard: k mod #GPUs (the rule snpCall applies to its -i path)."""
    m = re.search(r"best_split_(\d+)", split_name)
    k = int(m.group(1)) if m else 0
    return k % max(1, n_devices)


def splits_of_rank(n_splits, rank, world):
    """Static round-robin ownership of n_splits shards by `world` ranks (one process per GPU)."""
    return [k for k in range(n_splits) if k % world == rank]


def lpt_bins(weights, n_bins):
    """Greedy longest-processing-time binning, the rule of createOptimumSplit.py:56-60
    (heaviest genome first into the currently lightest bin). Returns bin index per item."""
    load = [0.0] * n_bins
    out = [0] * len(weights)
    for w, i in sorted(((w, i) for i, w in enumerate(weights)), reverse=True):
        b = load.index(min(load))
        load[b] += w
        out[i] = b
    return out


def reduce_over_ranks(dist, device, times, work):
    """times: per-rank durations -> max over ranks; work: per-rank amounts -> sum over ranks."""
    import torch
    if dist is None:
        return list(times), list(work)
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.tolist(), w.tolist()
