"""Worker of tests/test_gpu_multi.py: run under torchrun, one rank per GPU.  Each rank runs the MINORg-shaped class
on the config-1 data; rank 0 prints the failing gRNA as JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    data, out, mode = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from minorg_b200.MINORg import MINORg
    m = MINORg(directory=os.path.join(out, "rank%d" % rank), prefix="mg", device=local)
    m.target = os.path.join(data, "targets.fasta")
    m.ot_mismatch = 2
    if mode == "windows_gapped":  # general path (edit-distance prefilter + verifier) over window shards of one genome
        m.ot_gap = 2
    if mode == "files":          # three subject files over the ranks
        m.reference = {"TAIR10": os.path.join(data, "subset_ref_TAIR10.fasta")}
        m.add_query(os.path.join(data, "subset_9654.fasta"), alias="9654")
        m.add_query(os.path.join(data, "subset_9655.fasta"), alias="9655")
    else:                        # one genome split by window range
        m.reference = {"TAIR10": os.path.join(data, "subset_ref_TAIR10.fasta")}
    m.grna()
    m.filter_background()
    if rank == 0:
        fails = [s for s in m.grna_hits.seqs if m.grna_hits.get_gRNAseq_by_seq(s).check("background") is False]
        with open(os.path.join(out, "fails_%s.json" % mode), "w") as f:
            json.dump(fails, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
