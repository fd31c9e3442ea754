"""world_size-2 gloo worker for tests/test_host.py::test_multi_rank_sharding_gloo (CPU only)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import coracle                      # noqa: E402
from magot_b200 import engine, synth   # noqa: E402


def oracle_pass(contigs, tbl):
    """Stand-in for the device pass: payload text of every record via the C oracle (no framing)."""
    lens = np.array([len(c) for c in contigs])
    lo = np.clip(tbl.seg_start - 1, 0, lens[tbl.seg_contig])
    hi = np.clip(tbl.seg_end, 0, lens[tbl.seg_contig])
    text, off = coracle.splice(contigs, tbl.rec_seg_off, tbl.seg_contig, lo, hi, tbl.seg_strand)
    return text.tobytes()


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    layout = synth.contig_layout("insect", 300_000, 11)
    contigs = [a.tobytes() for a in synth.synth_genome_host(layout, 11)]
    ann = synth.synth_annotation(layout, 400, 12)
    tbl = ann.table("cds", framing=False)
    bounds = engine.shard_bounds(tbl.approx_bytes_per_record(), world)
    mine = oracle_pass(contigs, tbl.slice(bounds[rank], bounds[rank + 1]))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)          # host gather of the per-rank texts (test harness only)
    if rank == 0:
        whole = oracle_pass(contigs, tbl)
        ok = b"".join(gathered) == whole and all(len(g) > 0 for g in gathered)
        sizes = [len(g) for g in gathered]
        ok = ok and max(sizes) < 0.7 * sum(sizes)
        with open(sys.argv[1], "w") as fh:
            fh.write("OK" if ok else "MISMATCH %r" % sizes)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
