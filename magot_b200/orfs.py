"""Host wrapper of K4 (six-frame translation + ORF scan on device-resident contigs).

Replaces Sequence.get_orfs (genome.py:824-851) applied to whole contigs -- what dna2orfs
(genome_tools.py:145-180) intended -- with an optional minimum length (0 == reference).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import lib, check

ORF_DTYPE = np.dtype([("contig", np.int32), ("frame", np.int8), ("minus", np.int8), ("pad", np.int16),
                      ("start", np.int64), ("len", np.int64), ("aa_off", np.int64)])
assert ORF_DTYPE.itemsize == ctypes.sizeof(_lib.MgOrf)


def sixframe(device_genome, contig_lo, contig_hi, min_aa=0, stream=None):
    """ORFs of contigs [contig_lo, contig_hi) in reference order.
    Returns (records: structured array ORF_DTYPE, residues: bytes with the ORFs back to back)."""
    n_orf, n_bytes = ctypes.c_int64(0), ctypes.c_int64(0)
    check(lib.mg_sixframe_count(device_genome.handle, contig_lo, contig_hi, int(min_aa),
                                ctypes.byref(n_orf), ctypes.byref(n_bytes), stream))
    recs = np.zeros(n_orf.value, dtype=ORF_DTYPE)
    aa = np.empty(max(n_bytes.value, 1), dtype=np.uint8)
    check(lib.mg_sixframe_emit(device_genome.handle, ctypes.c_void_p(aa.ctypes.data),
                               ctypes.c_void_p(recs.ctypes.data) if n_orf.value else None, stream))
    check(lib.mg_stream_sync(device_genome.device, stream))
    return recs, aa[:n_bytes.value].tobytes()


def sixframe_list(device_genome, contig_ids, min_aa=0, stream=None):
    """ORFs of the listed contigs (any order), contig by contig in list order; same return as sixframe()."""
    ids = np.ascontiguousarray(contig_ids, dtype=np.int64)
    n_orf, n_bytes = ctypes.c_int64(0), ctypes.c_int64(0)
    check(lib.mg_sixframe_count_list(device_genome.handle, ids.size, ctypes.c_void_p(ids.ctypes.data), int(min_aa),
                                     ctypes.byref(n_orf), ctypes.byref(n_bytes), stream))
    recs = np.zeros(n_orf.value, dtype=ORF_DTYPE)
    aa = np.empty(max(n_bytes.value, 1), dtype=np.uint8)
    check(lib.mg_sixframe_emit(device_genome.handle, ctypes.c_void_p(aa.ctypes.data),
                               ctypes.c_void_p(recs.ctypes.data) if n_orf.value else None, stream))
    check(lib.mg_stream_sync(device_genome.device, stream))
    return recs, aa[:n_bytes.value].tobytes()


def lpt_shards(lengths, n_shards):
    """Longest-processing-time assignment of contigs to n_shards GPUs: shards[i] = contig indices (ascending) of GPU i."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    load = [0] * n_shards
    shards = [[] for _ in range(n_shards)]
    for c in order:
        k = min(range(n_shards), key=lambda i: load[i])
        shards[k].append(int(c))
        load[k] += int(lengths[c])
    return [sorted(x) for x in shards]


def contig_orfs(genome_sequence, seqid, min_aa=0):
    """(records, list of ORF strings) for one contig of a magot_b200.genome.GenomeSequence."""
    ci = genome_sequence.contig_index(seqid)
    recs, aa = sixframe(genome_sequence._engine().primary, ci, ci + 1, min_aa)
    text = aa.decode("latin-1")
    out = [text[int(r["aa_off"]):int(r["aa_off"]) + int(r["len"])] for r in recs]
    return recs, out
