"""Host-side multi-rank logic on CPU: world_size-2 gloo all-reduce of per-shard counts and the shard planners."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minorg_b200.distributed import allreduce_counts, merge_fail_flags, shard_files, shard_windows


def test_shard_planners_cover_once():
    for glen, L, world in [(1000, 20, 3), (64, 20, 8), (135_000_128, 23, 8), (10, 20, 2), (3_100_000_000, 20, 8)]:
        cuts = [shard_windows(glen, L, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == max(0, glen - L + 1)
        for (a, b), (c, d) in zip(cuts[:-1], cuts[1:]):
            assert b == c and a <= b and c <= d
    files = list("abcdefgh")
    for world in (1, 2, 3, 8):
        got = sorted(sum((shard_files(files, r, world) for r in range(world)), []))
        assert got == files


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    import random
    rng = random.Random(3)
    contigs = [("c%d" % i, "".join(rng.choice("ACGT") for _ in range(4000))) for i in range(4)]
    grnas = [contigs[i % 4][1][50 * i:50 * i + 20] for i in range(40)]
    mine = shard_files(list(range(4)), rank, world)
    local = np.zeros(40, dtype=np.int64)
    for c in mine:                                   # each rank screens its own genomes (oracle as a stand-in for the device)
        local += c_oracle.screen_hamming([contigs[c]], [], grnas, 2, threads=1)
    t = torch.from_numpy(np.concatenate([local, np.zeros(40, dtype=np.int64)]))
    allreduce_counts(t)
    if rank == 0:
        full = c_oracle.screen_hamming(contigs, [], grnas, 2, threads=2)
        q.put((t.numpy()[:40].tolist(), full.tolist(), merge_fail_flags(t.numpy(), 40).tolist()))
    dist.destroy_process_group()


def test_gloo_world2_allreduce_of_shard_counts():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, full, flags = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == full
    assert flags == [x > 0 for x in full]
