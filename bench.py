#!/usr/bin/env python
"""Benchmark of the hot path: exhaustive off-target screen, BASELINE.json metric "gRNA x Gbp screened/sec".

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the reference arm: CPU implementation)

Workload (BASELINE.json configs[1]): 8 synthetic genomes x 135 Mb (5 contigs of 27 Mb, i.i.d. ACGT),
64 planted 2-kb targets per genome (masked as on-targets), 10 000 SpCas9 20-nt gRNA enumerated from the
planted targets, ungapped <= 4 mismatches, both strands, PAM-less (the reference's defaults otherwise).
A step = one full screen of all gRNA against all 8 genomes (+ the all-reduce of per-gRNA counts for N>1).
The 8 genomes are divided over the N ranks (total work fixed => "strong" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_GENOMES = 8
CONTIGS_PER_GENOME = 5
TARGETS_PER_GENOME = 64
TARGET_LEN = 2000
GRNA_LEN = 20
MAX_MISMATCH = 4
# dram__bytes_read.sum + dram__bytes_write.sum of one full-size (8x135 Mb, N=1) scan launch, from the ncu capture
# summarised in profiles/ (see DESIGN.md); algorithmic bytes are 2 query chunks x 405 MB genome + 34 MB of query tables.
# The number is a CAPTURE, not measured in this run: the JSON names the file and the commit of the kernel it was taken on.
NCU_DRAM_BYTES_PER_LAUNCH = 860234752   # profiles/r01_pairs_v3_fullsize_dram.csv: 843.68 MB read + 16.56 MB written
NCU_DRAM_CAPTURE = {"file": "profiles/r01_pairs_v3_fullsize_dram.csv", "kernel_commit": "4de4244 (round 1; screen_pairs_kernel unchanged since)",
                    "metric": "dram__bytes_read.sum + dram__bytes_write.sum, one N=1 full-size launch"}
WINDOW_SHARD_HALO = 4096                 # chunk halo of the window-shard check: planted targets (masks) are 2 kb


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=135.0, help="size of each synthetic genome (Mb)")
    ap.add_argument("--grna", type=int, default=10000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the gapped / pattern side measurements (N=1 only)")
    ap.add_argument("--no-checks", action="store_true", help="skip the oracle parity slice and the window-shard check")
    ap.add_argument("--north-star-grna", type=int, default=1_000_000,
                    help="N=1 side measurement: gRNA count of the north-star criterion (<=4 mm / <=1 gap) on genome 0 alone (0 = skip)")
    ap.add_argument("--gapped", action="store_true",
                    help="NOT the BASELINE config: time the north star's <=4 mismatches / <=1 gap criterion on the same genomes "
                         "(edit-distance rows kernel + verifier); implies --no-cpu-baseline --no-other-modes")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- workload
def make_targets():
    """Planted target sequences of every genome (host, deterministic): {genome: [(name, seq str)]}."""
    out = {}
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for g in range(N_GENOMES):
        rng = np.random.Generator(np.random.PCG64(100 + g))
        codes = rng.integers(0, 4, size=(TARGETS_PER_GENOME, TARGET_LEN), dtype=np.uint8)
        out[g] = [("g%d_t%02d" % (g, t), lut[codes[t]].tobytes().decode()) for t in range(TARGETS_PER_GENOME)]
    return out


def genome_layout(genome_bases):
    clen = int(genome_bases) // CONTIGS_PER_GENOME
    return [clen] * CONTIGS_PER_GENOME


def plant_positions(g, clens):
    """(contig, start) of each planted target of genome g; non-overlapping by construction."""
    rng = np.random.Generator(np.random.PCG64(200 + g))
    pos = []
    per = TARGETS_PER_GENOME // len(clens) + 1
    for t in range(TARGETS_PER_GENOME):
        c = t % len(clens)
        slot = t // len(clens)
        width = clens[c] // per
        lo = slot * width
        start = lo + int(rng.integers(0, max(1, width - TARGET_LEN)))
        pos.append((c, start))
    return pos


def synth_genome_device(torch, g, clens, targets, device):
    """ASCII bytes of genome g on the device (uint8 tensor) with its targets planted."""
    gen = torch.Generator(device=device)
    gen.manual_seed(g)
    total = sum(clens)
    codes = torch.randint(0, 4, (total,), dtype=torch.uint8, device=device, generator=gen)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    ascii_ = lut[codes.long()] if total < (1 << 24) else torch.take(lut, codes.to(torch.int64))
    del codes
    starts = np.concatenate([[0], np.cumsum(clens)])
    for (c, s), (_, seq) in zip(plant_positions(g, clens), targets):
        t = torch.from_numpy(np.frombuffer(seq.encode(), dtype=np.uint8).copy()).to(device)
        a = int(starts[c]) + s
        ascii_[a:a + len(seq)] = t
    return ascii_


def choose_grnas(lib_mod, targets_all, n):
    """First n distinct NGG 20-mers enumerated (K2, device) from the planted targets; topped up with random."""
    from minorg_b200.pam import PAM
    from minorg_b200.fasta import reverse_complement
    prog = PAM(".NGG", GRNA_LEN).program()
    seen, out = set(), []
    for g in range(N_GENOMES):
        for _, seq in targets_all[g]:
            plus, minus = lib_mod.enumerate_sites(seq, prog, GRNA_LEN)
            rc = reverse_complement(seq)
            for text, starts in ((seq, plus), (rc, minus)):
                for s in starts.tolist():
                    k = text[s:s + GRNA_LEN]
                    if k not in seen:
                        seen.add(k)
                        out.append(k)
            if len(out) >= n:
                return out[:n]
    rng = np.random.Generator(np.random.PCG64(1000))
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    while len(out) < n:                               # in blocks: the north star's set is 10^6 gRNA
        blob = lut[rng.integers(0, 4, size=(n - len(out) + 64, GRNA_LEN))].tobytes().decode()
        for i in range(0, len(blob), GRNA_LEN):
            k = blob[i:i + GRNA_LEN]
            if k not in seen:
                seen.add(k)
                out.append(k)
                if len(out) >= n:
                    break
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------- CPU arm
def counts_checksum(c):
    """64-bit sum of counts[i] * (i + 1): equal at every N because the all-reduced count vector is."""
    c = np.asarray(c, dtype=np.uint64)
    w = np.arange(1, len(c) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int((c * w).sum(dtype=np.uint64))


def sass_counts(kernel):
    """Instruction counts of a shipped kernel's hot loop, from profiles/sass_counts.json (tools/sass_counts.py: cuobjdump
    of the library, committed with the excerpt it was counted on)."""
    with open(os.path.join(ROOT, "profiles", "sass_counts.json")) as f:
        return json.load(f)["kernels"][kernel]


def parity_slice(clens, seconds, rate_bases_per_s):
    """Window range of genome 0 for the oracle comparison: the end of contig 0 and the start of contig 1 - so the slice
    holds a contig boundary (windows hanging off both contig ends) - wide enough on one side to contain a whole planted,
    masked target, and as large as the CPU time budget allows.  -> (A, B): windows of contig 0 that start in
    [len0 - A, len0) and of contig 1 that start in [-L+1, B)."""
    pos = plant_positions(0, clens)
    end0 = [clens[0] - s for c, s in pos if c == 0]          # distance of a target's START from the end of contig 0
    start1 = [s + TARGET_LEN for c, s in pos if c == 1]      # END of a target of contig 1
    need_a = min(end0) + GRNA_LEN if end0 else None
    need_b = min(start1) + GRNA_LEN if start1 else None
    budget = int(max(200_000, rate_bases_per_s * seconds))
    a = b = budget // 2
    if need_a is not None and (need_b is None or need_a <= need_b):
        a = max(a, need_a)
        b = max(50_000, min(b, budget - a))
    elif need_b is not None:
        b = max(b, need_b)
        a = max(50_000, min(a, budget - b))
    return min(a, clens[0]), min(b, clens[1])


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    genome_bases = int(args.genome_mb * 1e6)
    clens = genome_layout(genome_bases)
    total_bases = sum(clens) * N_GENOMES
    if args.gapped:
        args.no_cpu_baseline = args.no_other_modes = True
    workload = ("synthetic %dx%.0f Mb genomes (5 contigs each, 64 planted+masked 2-kb targets), %d SpCas9 20-nt gRNA, "
                "%s <=%d mismatches, both strands, PAM-less" % (N_GENOMES, args.genome_mb, args.grna,
                                                               "<=1 gap and" if args.gapped else "ungapped", MAX_MISMATCH))
    config = {"workload": workload, "genomes": N_GENOMES, "genome_bp": sum(clens), "n_grna": args.grna,
              "max_mismatch": MAX_MISMATCH, "grna_len": GRNA_LEN, "cells_per_step": 2 * args.grna * total_bases,
              "l2": "inputs larger than L2 (packed genome %.0f MB per step)" % (total_bases * 0.375 / 1e6),
              "blastn": "unavailable in image (reference's BLASTn shell-out cannot run)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, clens, config)

    # ONE JSON line on stdout: native libraries write there too (NCCL prints its version line with NCCL_DEBUG=VERSION), so
    # the process-level stdout is pointed at stderr until the result line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj))
        sys.stdout.flush()

    import torch
    import torch.distributed as dist
    from minorg_b200 import _lib
    from minorg_b200.ot_regex import make_criteria

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    name, sms = _lib.init(local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- build the workload: this rank's genomes resident in HBM
    targets_all = make_targets()
    my_genomes = [g for g in range(N_GENOMES) if g % world == rank]
    grnas = choose_grnas(_lib, targets_all, args.grna)
    ascii_parts, names, mask_seqs = [], [], []
    for g in my_genomes:
        ascii_parts.append(synth_genome_device(torch, g, clens, targets_all[g], device))
        names += ["g%d_c%d" % (g, c) for c in range(len(clens))]
    for g in range(N_GENOMES):
        mask_seqs += [s for _, s in targets_all[g]]
    ascii_dev = torch.cat(ascii_parts) if len(ascii_parts) > 1 else ascii_parts[0]
    del ascii_parts
    offsets = np.concatenate([[0], np.cumsum(clens * len(my_genomes))]).astype(np.int64)
    genome = _lib.Genome.from_device_ascii(ascii_dev.data_ptr(), offsets, names, stream=stream)
    torch.cuda.synchronize()
    # on-target masks: K3 exact-occurrence finder over all planted targets (MINORg._mask_ontargets)
    found = genome.find_exact(mask_seqs, stream=stream)
    masks = sorted({(c, s, e) for _, c, s, e, _ in found})
    planted = sum(1 for g in my_genomes for _ in range(TARGETS_PER_GENOME))
    assert len(masks) >= planted, "mask finder missed planted targets (%d < %d)" % (len(masks), planted)
    genome.set_masks(masks, stream=stream)
    queries = _lib.Queries(grnas, stream=stream)
    crit = make_criteria(ot_mismatch=MAX_MISMATCH + 1, ot_gap=2 if args.gapped else 0, grna_length=GRNA_LEN)
    counts = torch.zeros(2 * len(grnas), dtype=torch.int32, device=device)

    def step():
        counts.zero_()
        _lib.screen(genome, queries, crit, None, counts.data_ptr(), stream=stream)
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = []
    ev[0].record()
    for i in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        counts.zero_()
        k0.record()
        _lib.screen(genome, queries, crit, None, counts.data_ptr(), stream=stream)
        k1.record()
        if world > 1:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        ev[i + 1].record()
        kev.append((k0, k1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = args.grna * total_bases / 1e9 / (ms_per_step / 1e3)

    # the all-reduced per-query counts of the last timed step: identical at every N (same total work), and so is
    # their checksum - SCALE runs prove multi-GPU correctness by printing one value at N = 1, 2, 4, 8
    n = len(grnas)
    steps_cnt, stage2 = _lib.last_screen_stats()
    sj_cand, sj_surv = _lib.last_seedjoin_stats() if args.gapped else (0, 0)      # of the last timed step (this rank)
    counts_host = counts.cpu().numpy().astype(np.uint32)
    host_counts = counts_host[:n].astype(np.int64) + counts_host[n:].astype(np.int64)
    fails = int((host_counts > 0).sum())
    checksum = counts_checksum(counts_host)

    # ---- checks that every bench run carries: counts against the CPU oracle on a slice of the full-size
    # genome, and one genome cut into chunk + halo shards against the same genome screened whole
    parity = None
    shard_check = None
    cpu_baseline = None
    threads = len(os.sched_getaffinity(0))
    if not args.no_checks:
        if rank == 0 and not args.gapped:
            want_baseline = world == 1 and not args.no_cpu_baseline
            parity, cpu_baseline = parity_check(torch, _lib, genome, queries, crit, grnas, masks, ascii_dev, clens,
                                                3.0, args.cpu_seconds if want_baseline else 0.0, threads, stream, device)
        if rank == 0 and args.gapped:
            parity = general_parity(torch, _lib, genome, queries, crit, grnas, masks, ascii_dev, clens, threads, stream, device,
                                    span=20000, n_sub=min(n, 256), ot_gap=2, max_mismatch=MAX_MISMATCH)
        shard_check = window_shard_check(torch, dist, _lib, rank, world, device, stream, queries, crit, grnas, clens,
                                         targets_all, ascii_dev if 0 in my_genomes else None, masks if 0 in my_genomes else None)

    # ---- side measurements (N=1): the other criteria of the path on the same resident genomes, one timed pass each
    # after one warm-up pass.  Not the headline: the gapped line is the north star's <=4 mm / <=1 gap criterion.
    other_modes = None
    if world == 1 and not args.no_other_modes:
        other_modes = []
        for label, n_sub, kw in (("gapped: <=4 mismatches / <=1 gap (edit-distance rows kernel + verifier)", min(n, 2048), dict(ot_mismatch=MAX_MISMATCH + 1, ot_gap=2)),
                                 ("pattern 0mg-10,1mg-11- with <=1 gap (BASELINE configs[2]; exact-piece kernel + verifier)", n, dict(ot_pattern="0mg-10,1mg-11-", ot_gap=2)),
                                 ("reference defaults: ot_mismatch=1 (exact off-target copies only)", n, dict(ot_mismatch=1))):
            q_sub = queries if n_sub == n else _lib.Queries(grnas[:n_sub], stream=stream)
            c_sub = torch.zeros(2 * n_sub, dtype=torch.int32, device=device)
            crit_m = make_criteria(grna_length=GRNA_LEN, **kw)
            _lib.screen(genome, q_sub, crit_m, None, c_sub.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            c_sub.zero_()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            _lib.screen(genome, q_sub, crit_m, None, c_sub.data_ptr(), stream=stream)
            m1.record()
            torch.cuda.synchronize()
            sec = m0.elapsed_time(m1) / 1e3
            entry = {"mode": label, "n_grna": n_sub, "seconds": sec, "value": n_sub * total_bases / 1e9 / sec,
                     "unit": "gRNA*Gbp/s", "tcell_per_s": 2.0 * n_sub * total_bases / sec / 1e12,
                     "grna_failed": int(((c_sub[:n_sub] + c_sub[n_sub:]) > 0).sum().item())}
            if not args.no_checks and ("ot_gap" in kw or "ot_pattern" in kw):
                # the same criterion on a slice of the full-size genome against the exhaustive C oracle (anchor counts)
                entry["parity_check"] = general_parity(torch, _lib, genome, q_sub, crit_m, grnas[:n_sub], masks, ascii_dev, clens, threads,
                                                       stream, device, span=20000 if "ot_pattern" not in kw else 200000,
                                                       n_sub=min(n_sub, 256), ot_gap=kw.get("ot_gap", 0),
                                                       max_mismatch=max(1, kw.get("ot_mismatch", 1)) - 1, pattern=kw.get("ot_pattern"))
            other_modes.append(entry)
            if q_sub is not queries:
                q_sub.free()
        if args.north_star_grna > 0:
            # the north star's own job, one GPU's share of it: 10^6 gRNA against ONE of the genomes, <=4 mm / <=1 gap (the
            # seed-and-join screen; on 8 GPUs each rank does exactly this: profiles/r02_bench_gapped_seedjoin_n8.json)
            n_ns = args.north_star_grna
            g_ns = choose_grnas(_lib, targets_all, n_ns)
            q_ns = _lib.Queries(g_ns, stream=stream)
            c_ns = torch.zeros(2 * n_ns, dtype=torch.int32, device=device)
            crit_ns = make_criteria(ot_mismatch=MAX_MISMATCH + 1, ot_gap=2, grna_length=GRNA_LEN)
            # genome 0's windows: from the padding before its first contig (anchors that hang off the contig start belong to
            # windows there) to half-way into the >= 64 invalid bases before the next genome's first contig
            w0 = 0
            w1 = genome.contig_global_start(len(clens)) - 32 if len(my_genomes) > 1 else -1
            _lib.screen(genome, q_ns, crit_ns, None, c_ns.data_ptr(), win_begin=w0, win_end=w0 + (1 << 20), stream=stream)   # builds the index
            torch.cuda.synchronize()
            c_ns.zero_()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            _lib.screen(genome, q_ns, crit_ns, None, c_ns.data_ptr(), win_begin=w0, win_end=w1, stream=stream)
            m1.record()
            torch.cuda.synchronize()
            sec = m0.elapsed_time(m1) / 1e3
            cand, surv = _lib.last_seedjoin_stats()
            entry = {"mode": "north star on one GPU: %d gRNA x genome 0 (%d bp), <=4 mismatches / <=1 gap (seed-and-join screen; the "
                             "8-GPU job is this on every rank)" % (n_ns, sum(clens)), "n_grna": n_ns, "seconds": sec,
                     "value": n_ns * sum(clens) / 1e9 / sec, "unit": "gRNA*Gbp/s", "tcell_per_s": 2.0 * n_ns * sum(clens) / sec / 1e12,
                     "candidates": cand, "stage2_candidates": surv, "candidates_per_cell": cand / (2.0 * n_ns * sum(clens)),
                     "anchors_total": int(c_ns.sum(dtype=torch.int64).item())}
            if not args.no_checks:
                entry["parity_check"] = general_parity(torch, _lib, genome, q_ns, crit_ns, g_ns, masks, ascii_dev, clens, threads, stream,
                                                       device, span=20000, n_sub=256, ot_gap=2, max_mismatch=MAX_MISMATCH)
            other_modes.append(entry)
            q_ns.free()
            del c_ns

    # ---- end-to-end through the host-buffer C-ABI call (H2D + pack + tables + screen + D2H)
    e2e = None
    if not args.no_e2e:
        host_ascii = torch.empty(ascii_dev.numel(), dtype=torch.uint8, pin_memory=True)
        host_ascii.copy_(ascii_dev)
        torch.cuda.synchronize()
        harr = host_ascii.numpy()
        gblob = np.frombuffer("".join(grnas).encode(), dtype=np.uint8)
        del ascii_dev
        genome.free()
        torch.cuda.empty_cache()
        red = torch.zeros(n, dtype=torch.int32, device=device)

        def e2e_step():
            c = _lib.screen_host(harr, offsets, masks, gblob, n, GRNA_LEN, crit, None, stream=stream)
            if world > 1:
                red.copy_(torch.from_numpy(c.astype(np.int32)))
                dist.all_reduce(red, op=dist.ReduceOp.SUM)
                return red.cpu().numpy()
            return c
        for _ in range(min(args.warmup, 1)):
            e2e_step()
        barrier()
        e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            c = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e_ms = dt.item() * 1e3 / e_steps
        # at every N: the (reduced) counts of the host-buffer path equal the resident path's
        e2e_equal = bool(np.array_equal(np.asarray(c, dtype=np.int64), host_counts))
        assert e2e_equal, "end-to-end counts differ from the resident-genome counts (world %d)" % world
        e2e = {"value": args.grna * total_bases / 1e9 / (e_ms / 1e3), "unit": "gRNA*Gbp/s", "ms_per_step": e_ms,
               "steps": e_steps, "h2d_bytes_per_step": int(len(harr) * world + gblob.nbytes * world),
               "d2h_bytes_per_step": int(8 * n * world), "counts_equal_resident_path": e2e_equal}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    # The scan is bound by the shared-memory load pipe (LSU) and by the ALU pipe (LOP3) at nearly the same level; HBM
    # carries 0.375 B/base once per query chunk and is ~4 orders of magnitude below its roofline.  Denominators:
    # mg_microbench measures both pipes on this GPU.  Work per (warp, 32-query group) step: instruction counts read from
    # the SASS of the shipped kernel (profiles/sass_counts.json, tools/sass_counts.py) x the fraction of steps that go
    # past the vote, counted by the kernel itself (mg_last_screen_stats).
    peaks, peaks_src = measured_peaks()
    lop3 = _lib.microbench(0, stream)     # Glane-ops/s, measured here
    lds = _lib.microbench(1, stream)
    imad = _lib.microbench(7, stream)
    both = _lib.microbench(8, stream)
    int_peaks = {"lop3_glane_ops": lop3, "lds32_glane_ops": lds, "imad_glane_ops": imad, "lop3_imad_interleaved_glane_ops": both}
    my_bases = sum(clens) * len(my_genomes)
    cells_per_launch = 2.0 * n * my_bases
    achieved = cells_per_launch / (kern_ms / 1e3) / 1e12
    if args.gapped and sj_cand:
        # seed-and-join (sj_scan_kernel): the unit of work is a CANDIDATE (an index entry compared with the text), counted by
        # the kernel itself; the ceiling is one warp instruction per scheduler and cycle over the warp instructions the kernel
        # executes per 32 candidates (ncu, profiles/r02_seedjoin_ncu.json: probes, queueing and both stages included)
        f_hz = (clocks or {}).get("sm_mhz") or 1965.0
        try:
            with open(os.path.join(ROOT, "profiles", "r02_seedjoin_ncu.json")) as fh:
                prof = json.load(fh)
            per32 = prof["warp_instructions"] / (prof["candidates"] / 32.0)
        except (OSError, KeyError, ValueError):
            prof, per32 = None, None
        cand_s = sj_cand / (kern_ms / 1e3) / 1e9
        peak_c = sms * 4 * f_hz * 1e6 * 32.0 / per32 / 1e9 if per32 else None
        roofline = {"bound": "issue slots", "achieved": cand_s, "peak": peak_c, "unit": "Gcandidate/s",
                    "frac": cand_s / peak_c if peak_c else None, "traffic": prof.get("dram_bytes_per_launch") if prof else None,
                    "candidates_per_launch": sj_cand, "stage2_candidates_per_launch": sj_surv,
                    "candidates_per_cell": sj_cand / cells_per_launch,
                    "effective_tcell_per_s": achieved,
                    "warp_instructions_per_32_candidates": per32,
                    "peak_source": "4 schedulers x %d SMs x %.0f MHz (sampled during the run), one warp instruction each per cycle, over the "
                                   "kernel's own instruction count per candidate (ncu capture in profiles/)" % (sms, f_hz),
                    "kernel": "sj_scan_kernel (+ the edit-distance rows kernel and verifier on the windows near N runs, contig ends "
                              "and masks)", "kernel_ms": kern_ms,
                    "note": "the screen stays exhaustive: every (gRNA, anchor) that satisfies the criterion is found through the "
                            "pigeonhole routes of DESIGN.md 4; counts equal the DP path's and the C oracle's (parity_check, tests)"}
        out = {"metric": "gRNA x Gbp screened/sec (exhaustive, device-timed)", "value": value, "unit": "gRNA*Gbp/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 (bit planes of 2-bit bases)",
               "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": 4 * args.steps * world,
               "roofline": roofline, "cpu_baseline": None, "device": name, "int_peaks": int_peaks,
               "cells_per_s": 2.0 * args.grna * total_bases / (ms_per_step / 1e3), "grna_failed": fails,
               "counts_checksum": "0x%016x" % checksum, "counts_total": int(host_counts.sum()),
               "parity_check": parity, "window_shard_check": shard_check,
               "side_measurement": "north-star criterion (<=4 mm / <=1 gap), not BASELINE.json's bench config"}
        emit(out)
        if world > 1:
            dist.destroy_process_group()
        return 0
    if args.gapped:
        # gapped_rows_kernel<20,20>: issue slots per (32 gRNA, text character) from the SASS of the per-character loop, less
        # the cold survivor path; one slot per scheduler and cycle = the measured LOP3+IMAD interleaved rate
        sc = sass_counts("gapped_rows_kernel<20,20>")["per_character"]
        slots = sc.get("LOP3", 0) + sc.get("IMAD", 0) + sc.get("LDS", 0) + sc.get("SHF", 0) + 4      # + loop control
        f_hz = (clocks or {}).get("sm_mhz") or 1965.0
        peak_rows = sms * 4 * f_hz * 1e6 * 32.0 * 32.0 / slots / 1e12      # one instruction per scheduler and cycle
        roofline = {"bound": "issue slots (ALU + FMA pipes together)", "achieved": achieved,
                    "peak": peak_rows, "unit": "Tcell/s", "frac": achieved / peak_rows, "traffic": None,
                    "issue_slots_per_character_and_group": slots, "sass": "profiles/r02_sass_gapped_rows_L20_char_loop.txt",
                    "peak_source": "4 schedulers x %d SMs x %.0f MHz (sampled during the run), one warp instruction each per cycle; "
                                   "a 1:1 LOP3/IMAD microbenchmark (mg_microbench 8) sustains only %.0f Glane-op/s = %.2f of that "
                                   "ceiling on this GPU, the kernel's own mix more" % (sms, f_hz, both, both * 1e9 / (sms * 4 * f_hz * 1e6 * 32.0)),
                    "kernel": "gapped_rows_kernel<20,20> + verify_kernel", "kernel_ms": kern_ms,
                    "note": "achieved includes the verifier pass over the edit-distance survivors (8e-5 of the cells)"}
        out = {"metric": "gRNA x Gbp screened/sec (exhaustive, device-timed)", "value": value, "unit": "gRNA*Gbp/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 (bit-sliced 2-bit bases)",
               "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": None,
               "roofline": roofline, "cpu_baseline": None, "device": name, "int_peaks": int_peaks,
               "cells_per_s": 2.0 * args.grna * total_bases / (ms_per_step / 1e3), "grna_failed": fails,
               "counts_checksum": "0x%016x" % checksum, "counts_total": int(host_counts.sum()),
               "parity_check": parity, "window_shard_check": shard_check,
               "side_measurement": "north-star criterion (<=4 mm / <=1 gap), not BASELINE.json's bench config"}
        emit(out)
        if world > 1:
            dist.destroy_process_group()
        return 0
    kname = "screen_pairs_kernel<%d,%d,1024,0>" % (GRNA_LEN, MAX_MISMATCH)
    sc = sass_counts(kname)
    s1, s2 = sc["stage1_hot_path"], sc["stage2_fall_through"]
    f2 = (stage2 / steps_cnt) if steps_cnt else 1.0
    loads_per_step = s1["LDS"] + f2 * s2.get("LDG", 0)
    alu2 = sum(v for k, v in s2.items() if k in ("LOP3", "IADD3", "VIADD", "SHF", "PRMT", "ISETP", "MOV"))
    alu_per_step = s1["LOP3"] + f2 * alu2
    peak_lsu = lds * 1e9 * 32.0 / loads_per_step / 1e12
    peak_alu = lop3 * 1e9 * 32.0 / alu_per_step / 1e12
    peak = min(peak_lsu, peak_alu)
    brute = lds * 1e9 * 32.0 / GRNA_LEN / 1e12
    n_chunks = -(-((2 * n + 31) // 32) // 352)     # 352 query groups of pair tables fit one SM's shared memory
    alg_bytes = my_bases * 0.375 * n_chunks   # packed genome (2-bit bases + 1-bit invalid plane) streamed once per query chunk
    roofline = {"bound": "lsu(shared-memory loads)" if peak_lsu <= peak_alu else "alu(lop3)", "achieved": achieved,
                "peak": peak, "unit": "Tcell/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (args.genome_mb == 135.0 and world == 1) else None,
                "traffic_source": NCU_DRAM_CAPTURE,
                "peak_source": "mg_microbench on this GPU: LOP3 %.0f Glane-op/s, LDS.32 %.0f Glane-op/s" % (lop3, lds),
                "work_per_step": {"second_stage_fraction": f2, "lds_per_lane": loads_per_step, "alu_per_lane": alu_per_step,
                                  "source": "profiles/sass_counts.json (cuobjdump of the shipped library): stage 1 %d LDS + %d LOP3, "
                                            "stage 2 %d LDG + %d ALU-pipe instructions" % (s1["LDS"], s1["LOP3"], s2.get("LDG", 0), alu2)},
                "ceilings_tcell_s": {"lsu_shared": peak_lsu, "alu_lop3": peak_alu, "lsu_without_early_out": brute},
                "frac_of_ceiling_without_early_out": achieved / brute,
                "kernel": kname, "kernel_ms": kern_ms,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                        "achieved_gbs": alg_bytes / (kern_ms / 1e3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "peak_source": peaks_src,
                        "note": "0.375 B/base once per query chunk: HBM is ~3 orders below its roofline; the scan is integer-pipe bound"}}

    out = {"metric": "gRNA x Gbp screened/sec (exhaustive, device-timed)", "value": value, "unit": "gRNA*Gbp/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 (bit-sliced 2-bit bases)",
           "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * world,
           "roofline": roofline, "cpu_baseline": cpu_baseline, "device": name, "int_peaks": int_peaks,
           "cells_per_s": 2.0 * args.grna * total_bases / (ms_per_step / 1e3), "grna_failed": fails,
           "counts_checksum": "0x%016x" % checksum, "counts_total": int(host_counts.sum()),
           "parity_check": parity, "window_shard_check": shard_check,
           "other_modes": other_modes}
    emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


def general_parity(torch, _lib, genome, q, crit, grnas, masks, ascii_dev, clens, threads, stream, device, span, n_sub, ot_gap=0,
                   max_mismatch=None, pattern=None):
    """Gapped / pattern criteria on the FULL-SIZE resident genome against the exhaustive C oracles (oracle_screen_nopattern_win
    / oracle_screen_pattern) on the windows that start within `span` bases of the contig 0 / contig 1 boundary of genome 0, for the
    first n_sub gRNA: per-oriented-query anchor counts, one to one."""
    from oracle import c_oracle
    c_oracle.build()
    len0, len1 = clens[0], clens[1]
    span = int(min(span, len0, len1))
    n = len(grnas)
    host = ascii_dev[:len0 + len1].cpu().numpy()
    contigs = [("c0", host[:len0].tobytes()), ("c1", host[len0:len0 + len1].tobytes())]
    m01 = [(c, s, e) for c, s, e in masks if c < 2]
    ranges = [(len0 - span, len0), (-GRNA_LEN, span)]
    t0 = time.perf_counter()
    if pattern is None:
        want = c_oracle.screen_nopattern(contigs, m01, grnas[:n_sub], max_mismatch, max(1, ot_gap) - 1, threads=threads, win_ranges=ranges)
    else:
        want = c_oracle.screen_pattern(contigs, m01, grnas[:n_sub], pattern, ot_gap, threads=threads, win_ranges=ranges)
    cpu_s = time.perf_counter() - t0
    want = want.astype(np.int64)
    c = torch.zeros(2 * n, dtype=torch.int32, device=device)
    _lib.screen(genome, q, crit, None, c.data_ptr(), win_begin=genome.contig_global_start(0) + len0 - span,
                win_end=genome.contig_global_start(1) + span, stream=stream)
    torch.cuda.synchronize()
    got = c.cpu().numpy().astype(np.int64)
    got = np.concatenate([got[:n_sub], got[n:n + n_sub]])
    equal = bool(np.array_equal(got, want))
    out = {"window_bp": 2 * span + GRNA_LEN - 1, "n_grna": n_sub, "equal": equal, "anchors_total": int(want.sum()),
           "queries_with_hits": int((want > 0).sum()),
           "oracle": "oracle/screen_oracle.c %s, %d threads, %.1f s" % ("oracle_screen_pattern" if pattern else "oracle_screen_nopattern_win", threads, cpu_s)}
    if not equal:
        bad = np.nonzero(got != want)[0]
        out["first_differences"] = [(int(i), int(got[i]), int(want[i])) for i in bad[:5]]
    assert equal, "parity: device anchor counts differ from the C oracle: %s" % out
    return out


def parity_check(torch, _lib, genome, queries, crit, grnas, masks, ascii_dev, clens, seconds, baseline_seconds, threads, stream, device):
    """Device counts on a window range of the FULL-SIZE resident genome 0 against oracle/screen_oracle.c on the same
    windows: the end of contig 0 + the start of contig 1 (a contig boundary with windows hanging off both ends) reaching
    past a planted, masked target that the gRNA were cut from.  With baseline_seconds > 0 the bit-sliced CPU port (the
    fastest CPU implementation in the repo, same counts) is then timed on a larger slice of the same kind = cpu_baseline,
    and its counts are compared with the device's too.  -> (parity_check dict, cpu_baseline dict or None)."""
    from oracle import c_oracle
    c_oracle.build()
    n = len(grnas)
    len0, len1 = clens[0], clens[1]
    host = ascii_dev[:len0 + len1].cpu().numpy()
    off = np.array([0, len0, len0 + len1], dtype=np.int64)
    m01 = [(c, s, e) for c, s, e in masks if c < 2]
    probe = min(len0, 40000)
    t0 = time.perf_counter()
    c_oracle.screen_hamming_arrays(host, off, m01, grnas, MAX_MISMATCH, threads, [(0, probe), (0, 0)])
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    a, b = parity_slice(clens, seconds, rate)
    ranges = [(len0 - a, len0), (-GRNA_LEN, b)]
    t0 = time.perf_counter()
    want = c_oracle.screen_hamming_arrays(host, off, m01, grnas, MAX_MISMATCH, threads, ranges).astype(np.int64)
    cpu_s = time.perf_counter() - t0
    wb = genome.contig_global_start(0) + len0 - a
    we = genome.contig_global_start(1) + b
    c = torch.zeros(2 * n, dtype=torch.int32, device=device)
    _lib.screen(genome, queries, crit, None, c.data_ptr(), win_begin=wb, win_end=we, stream=stream)
    torch.cuda.synchronize()
    got = (c[:n] + c[n:]).cpu().numpy().astype(np.int64)
    equal = bool(np.array_equal(got, want))
    in_slice = [(c_, s, e) for c_, s, e in m01 if (c_ == 0 and s >= len0 - a) or (c_ == 1 and e <= b)]
    window_bp = a + b + GRNA_LEN - 1
    parity = {"window_bp": int(window_bp), "n_grna": n, "equal": equal,
              "windows": "genome 0: contig 0 starts [%d, %d) + contig 1 starts [%d, %d): one contig boundary" % (len0 - a, len0, -GRNA_LEN + 1, b),
              "masked_targets_in_slice": len(in_slice), "max_mismatch": MAX_MISMATCH,
              "counts_total": int(want.sum()), "grna_with_hits": int((want > 0).sum()),
              "oracle": "oracle/screen_oracle.c oracle_screen_hamming_win, %d threads, %.1f s" % (threads, cpu_s)}
    if not equal:
        bad = np.nonzero(got != want)[0]
        parity["first_differences"] = [(int(i), int(got[i]), int(want[i])) for i in bad[:5]]
    assert equal, "parity_check: device counts differ from the CPU oracle: %s" % parity.get("first_differences")
    parity["scalar_oracle_grna_gbp_per_s"] = n * window_bp / 1e9 / cpu_s
    cpu = None
    if baseline_seconds > 0:
        t0 = time.perf_counter()
        c_oracle.screen_hamming_arrays(host, off, m01, grnas, MAX_MISMATCH, threads, [(0, probe), (0, 0)], bitsliced=True)
        rate2 = probe / max(time.perf_counter() - t0, 1e-6)
        a2, b2 = parity_slice(clens, baseline_seconds, rate2)
        ranges2 = [(len0 - a2, len0), (-GRNA_LEN, b2)]
        t0 = time.perf_counter()
        want2 = c_oracle.screen_hamming_arrays(host, off, m01, grnas, MAX_MISMATCH, threads, ranges2, bitsliced=True).astype(np.int64)
        cpu2_s = time.perf_counter() - t0
        c.zero_()
        _lib.screen(genome, queries, crit, None, c.data_ptr(), win_begin=genome.contig_global_start(0) + len0 - a2,
                    win_end=genome.contig_global_start(1) + b2, stream=stream)
        torch.cuda.synchronize()
        got2 = (c[:n] + c[n:]).cpu().numpy().astype(np.int64)
        bp2 = a2 + b2 + GRNA_LEN - 1
        equal2 = bool(np.array_equal(got2, want2))
        parity["bitsliced_port"] = {"window_bp": int(bp2), "equal": equal2, "counts_total": int(want2.sum())}
        assert equal2, "parity_check: device counts differ from the bit-sliced CPU port"
        cpu = {"value": n * bp2 / 1e9 / cpu2_s, "unit": "gRNA*Gbp/s", "cores": threads, "kind": "port",
               "implementation": "oracle_screen_hamming_bitsliced: the device's pair-word algorithm on 64-bit words (plain C, pthreads); "
                                 "the scalar XOR/popcount oracle runs at %.1f" % (n * window_bp / 1e9 / cpu_s),
               "sample": "%d gRNA x %d bp of genome 0 around the contig 0/1 boundary (masks applied), <=%d mismatches, both strands, %.1f s; "
                         "counts equal the device's on the same windows" % (n, bp2, MAX_MISMATCH, cpu2_s)}
    return parity, cpu


def window_shard_check(torch, dist, _lib, rank, world, device, stream, queries, crit, grnas, clens, targets_all, ascii0, masks_full):
    """Genome 0 cut into chunk + halo shards (mg_genome_from_device_ascii_chunk; one per rank, or 4 on one GPU when N=1):
    every shard packs only its own bases, finds the planted targets that START in its range, the mask lists are gathered,
    every shard screens the windows it owns, counts are summed (all-reduce) and compared on rank 0 with genome 0
    screened whole.  This is how BASELINE configs[4] (one 3.1 Gb genome over 8 GPUs) runs."""
    from minorg_b200.distributed import shard_windows
    n = len(grnas)
    parts = world if world > 1 else 4
    if ascii0 is None:
        ascii0 = synth_genome_device(torch, 0, clens, targets_all[0], device)
    else:
        ascii0 = ascii0[:sum(clens)]
    off0 = np.concatenate([[0], np.cumsum(clens)]).astype(np.int64)
    names0 = ["g0_c%d" % c for c in range(len(clens))]
    glen0 = _lib.layout_global_length(off0)
    mask_seqs0 = [sq for _, sq in targets_all[0]]
    mine = list(range(parts)) if world == 1 else [rank]
    chunks, found = [], []
    for r in mine:
        wb, we = shard_windows(glen0, GRNA_LEN, r, parts)
        g = _lib.Genome.from_device_ascii_chunk(ascii0.data_ptr(), off0, names0, wb, we, halo=WINDOW_SHARD_HALO, stream=stream)
        found += [(c, s, e) for _, c, s, e, _ in g.find_exact(mask_seqs0, stream=stream)]
        chunks.append(g)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, found)
        found = [m for part in gathered for m in part]
    masks0 = sorted(set(found))
    total = torch.zeros(2 * n, dtype=torch.int32, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for g in chunks:
        g.set_masks(masks0, stream=stream)
    torch.cuda.synchronize()
    e0.record()
    for g in chunks:
        _lib.screen(g, queries, crit, None, total.data_ptr(), stream=stream)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    chunk_bytes = [g.device_bytes for g in chunks]
    for g in chunks:
        g.free()
    if rank != 0:
        return None
    whole = _lib.Genome.from_device_ascii(ascii0.data_ptr(), off0, names0, stream=stream)
    ref_masks = sorted({(c, s, e) for c, s, e in masks_full if c < len(clens)})
    whole.set_masks(ref_masks, stream=stream)
    ref = torch.zeros(2 * n, dtype=torch.int32, device=device)
    _lib.screen(whole, queries, crit, None, ref.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    equal = bool(torch.equal(ref, total))
    masks_equal = masks0 == ref_masks
    out = {"genome": "genome 0 (%d bp, %d contigs)" % (sum(clens), len(clens)), "shards": parts,
           "ranks": world, "halo_bp": WINDOW_SHARD_HALO, "equal": equal, "masks_equal": masks_equal,
           "masks_found_by_shards": len(masks0), "ms": ms, "counts_total": int(ref.sum().item()),
           "shard_device_bytes": int(max(chunk_bytes)), "whole_genome_device_bytes": int(whole.device_bytes)}
    whole.free()
    assert equal and masks_equal, "window-shard check failed: %s" % out
    return out


def reference_arm(args, clens, config):
    """The reference's own path cannot run here (blastn, Biopython and pyfaidx are absent from the image and
    there is no network), so - as the task's tier framing prescribes - this arm times the CPU restatement of
    the reference's criteria (oracle/screen_oracle.c) on all host threads, on a bounded sample per step."""
    threads = len(os.sched_getaffinity(0))
    from minorg_b200.fasta import reverse_complement
    targets_all = make_targets()
    # gRNA: same construction as the GPU arm but with the oracle's enumerator semantics (NGG sites, both strands)
    seen, grnas = set(), []
    for g in range(N_GENOMES):
        for _, seq in targets_all[g]:
            rc = reverse_complement(seq)
            for text in (seq, rc):
                for s in range(len(text) - GRNA_LEN - 2):
                    if text[s + GRNA_LEN + 1:s + GRNA_LEN + 3] == "GG":
                        k = text[s:s + GRNA_LEN]
                        if k not in seen:
                            seen.add(k)
                            grnas.append(k)
            if len(grnas) >= args.grna:
                break
        if len(grnas) >= args.grna:
            break
    grnas = grnas[:args.grna]
    rng = np.random.Generator(np.random.PCG64(0))
    nb_max = min(sum(clens), 40_000_000)
    ascii_host = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=nb_max, dtype=np.uint8)]
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    from oracle import c_oracle
    c_oracle.build()
    probe = min(nb_max, max(20000, 4000 * threads))
    t0 = time.perf_counter()
    c_oracle.screen_hamming_arrays(ascii_host[:probe], np.array([0, probe], dtype=np.int64), [], grnas, MAX_MISMATCH, threads, bitsliced=True)
    rate = probe / (time.perf_counter() - t0)
    nb = int(min(nb_max, max(probe, rate * per_step)))
    off = np.array([0, nb], dtype=np.int64)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_oracle.screen_hamming_arrays(ascii_host[:nb], off, [], grnas, MAX_MISMATCH, threads, bitsliced=True)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    value = len(grnas) * nb / 1e9 / sec
    sample = ("%d gRNA x %d bp per step, %d threads; NOT the GPU arm's inputs: an independent uniform random sample (seed 0) "
              "with no planted targets and no masks, gRNA enumerated by the oracle's own NGG scan" % (len(grnas), nb, threads))
    out = {"impl": "reference", "metric": "gRNA x Gbp screened/sec (exhaustive, device-timed)", "value": value,
           "unit": "gRNA*Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "u64 (bit-sliced pair words, 64 queries per word)", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": value, "unit": "gRNA*Gbp/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "gRNA*Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
