"""Multi-GPU plumbing for the screen: one process per GPU (torch.distributed, NCCL over NVLink), work split by
genome file or - for one large genome - by chunk (+ halo) of the window range, and ONE all-reduce of the per-gRNA counts.

The reference's only parallelism is a process pool over genome FILES (functions.py:76-100, MINORg.py:2182-2195);
cells are independent, so the only exchange step is the reduction of per-gRNA results."""
import numpy as np


def shard_files(items, rank, world):
    """Round-robin assignment of subject files (or genomes) to ranks; every item is owned exactly once."""
    return [x for i, x in enumerate(items) if i % world == rank]


def shard_windows(global_length, grna_len, rank, world):
    """Window-start range [begin, end) owned by `rank` when ONE subject is split by chunk.  A window belongs to the
    shard holding its first base.  Each rank packs only its range plus a halo (mg_genome_from_ascii_chunk: gRNA length +
    gaps + PAM flank, or the longest mask sequence for the mask finder), so nothing is exchanged but the per-gRNA counts;
    global_length comes from _lib.layout_global_length(offsets) before anything is on a device.  (With the whole
    packed genome resident on every GPU the same ranges work through mg_screen's win_begin/win_end.)"""
    n_windows = max(0, int(global_length) - int(grna_len) + 1)
    per = -(-n_windows // world)
    per = -(-per // 32) * 32                      # warp-aligned cuts
    begin = min(n_windows, rank * per)
    end = min(n_windows, begin + per)
    return begin, end


def allreduce_counts(counts):
    """Sum per-query counts over ranks in place (torch tensor on the device for NCCL, on the CPU for gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def merge_fail_flags(counts_fwd_rev, n):
    """counts of the 2N oriented queries -> boolean fail flag per gRNA."""
    c = np.asarray(counts_fwd_rev, dtype=np.int64)
    return (c[:n] + c[n:2 * n]) > 0
