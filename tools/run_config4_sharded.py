#!/usr/bin/env python
"""BASELINE configs[4]: ONE synthetic 3.1 Gb human-scale genome (24 contigs, seed 30), <= 4 mismatches, both strands,
chunk + halo sharded over the ranks of a torch.distributed job (north_star "genomes are sharded by chunk (with
gRNA-length halo)"): every rank synthesises and packs ONLY the contigs its chunk touches (mg_genome_from_device_ascii_chunk
stores chunk + halo), screens the windows it owns and the per-gRNA counts are summed with one NCCL all-reduce.

    python tools/run_config4_sharded.py --grna 20000                      (one GPU: the whole genome, reference checksum)
    python -m torch.distributed.run --nproc-per-node 8 ... tools/run_config4_sharded.py --grna 20000 --grna 1000000

The counts checksum of a sharded run must equal the one-GPU run's at the same gRNA count.  One JSON line per gRNA count."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from minorg_b200 import _lib  # noqa: E402
from minorg_b200.distributed import shard_windows  # noqa: E402
from minorg_b200.ot_regex import make_criteria  # noqa: E402

GENOME_BP, N_CONTIGS, SEED, L, HALO = 3_100_000_000, 24, 30, 20, 256


def contig_ascii(dev, c, n):
    """Contig c is generated from its own seed, so that any rank can make exactly the contigs it needs."""
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED * 1000 + c)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    return torch.take(lut, torch.randint(0, 4, (n,), dtype=torch.int64, device=dev, generator=gen))


def checksum(c):
    c = np.asarray(c, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int((c * np.arange(1, len(c) + 1, dtype=np.uint64)).sum(dtype=np.uint64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grna", type=int, action="append", default=None)
    ap.add_argument("--genome-bp", type=int, default=GENOME_BP)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    name, sms = _lib.init(local)
    stream = torch.cuda.current_stream().cuda_stream
    clen = args.genome_bp // N_CONTIGS
    clens = [clen] * (N_CONTIGS - 1) + [args.genome_bp - clen * (N_CONTIGS - 1)]
    offsets = np.concatenate([[0], np.cumsum(clens)]).astype(np.int64)
    glen = _lib.layout_global_length(offsets)
    begin, end = shard_windows(glen, L, rank, world)
    # the contigs whose global span [cstart, cend) meets [begin - HALO, end + HALO): global starts follow the library's
    # layout (64 invalid bases of padding, starts aligned to 32) - recomputed here only to decide what to synthesise
    pos, cstart = 64, []
    for n in clens:
        pos = (pos + 31) // 32 * 32
        cstart.append(pos)
        pos += n + 64
    need = [c for c in range(N_CONTIGS) if cstart[c] < end + HALO and cstart[c] + clens[c] > begin - HALO]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    buf = torch.empty(int(offsets[need[-1] + 1] - offsets[need[0]]), dtype=torch.uint8, device=dev)
    for c in need:
        a = int(offsets[c] - offsets[need[0]])
        buf[a:a + clens[c]] = contig_ascii(dev, c, clens[c])
    # the library indexes the ASCII of contig c at d_seq + offsets[c]: hand it the address the whole genome WOULD start at
    base_ptr = buf.data_ptr() - int(offsets[need[0]])
    genome = _lib.Genome.from_device_ascii_chunk(base_ptr, offsets, ["chr%d" % (c + 1) for c in range(N_CONTIGS)], begin, end,
                                                 halo=HALO, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    setup_ms = e0.elapsed_time(e1)
    ascii_bytes = int(buf.numel())
    del buf
    torch.cuda.empty_cache()
    crit = make_criteria(ot_mismatch=5, ot_gap=0, grna_length=L)
    rng = np.random.Generator(np.random.PCG64(1000))
    all_grna = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(max(args.grna or [20000]), L))]
    for n in (args.grna or [20000]):
        grnas = [bytes(r).decode() for r in all_grna[:n]]
        q = _lib.Queries(grnas, stream=stream)
        counts = torch.zeros(2 * n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        _lib.screen(genome, q, crit, None, counts.data_ptr(), stream=stream)
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        t1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sec = ms.item() / 1e3
        if rank == 0:
            c = counts.cpu().numpy().astype(np.uint32)
            print(json.dumps({"config": "configs[4]: one %.2f Gb genome (24 contigs), <=4 mm both strands, chunk + halo sharded" % (args.genome_bp / 1e9),
                              "n_gpus": world, "n_grna": n, "seconds": sec, "grna_gbp_per_s": n * args.genome_bp / 1e9 / sec,
                              "tcell_per_s": 2.0 * n * args.genome_bp / sec / 1e12, "counts_checksum": "0x%016x" % checksum(c),
                              "counts_total": int(c.astype(np.int64).sum()), "grna_failed": int(((c[:n].astype(np.int64) + c[n:]) > 0).sum()),
                              "rank0_chunk": [begin, end], "rank0_contigs_synthesised": len(need), "rank0_ascii_bytes": ascii_bytes,
                              "rank0_device_bytes": genome.device_bytes, "whole_genome_device_bytes": int(glen * 0.375),
                              "rank0_setup_ms": setup_ms, "device": name,
                              "rank0_seedjoin_candidates": _lib.last_seedjoin_stats()[0],      # > 0: the seed-and-join screen ran (>= 10^5 gRNA)
                              "MG_SEEDJOIN": os.environ.get("MG_SEEDJOIN", "")}))
            sys.stdout.flush()
        q.free()
    genome.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
