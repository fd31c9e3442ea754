"""SoA interval tables (pure numpy; no CUDA library needed to build or slice them).

`RecordTable` is what the host flattener produces and `mg_plan_create` (include/magot_b200.h) consumes: per output
record a list of segments (contig id, start, end, strand) in the reference's emission order
(ParentAnnotation.get_fasta, genome.py:687-705) plus the literal framing of the FASTA record.
"""
import numpy as np


class RecordTable(object):
    """SoA tables for n_rec output records (see mg_plan_create in include/magot_b200.h)."""

    __slots__ = ("rec_seg_off", "seg_contig", "seg_start", "seg_end", "seg_strand", "rec_lit_off",
                 "rec_pre_len", "rec_suf_len", "lit", "rec_phase")

    def __init__(self, rec_seg_off, seg_contig, seg_start, seg_end, seg_strand, rec_lit_off, rec_pre_len,
                 rec_suf_len, lit, rec_phase=None):
        self.rec_seg_off = np.ascontiguousarray(rec_seg_off, dtype=np.int64)
        self.seg_contig = np.ascontiguousarray(seg_contig, dtype=np.int32)
        self.seg_start = np.ascontiguousarray(seg_start, dtype=np.int64)
        self.seg_end = np.ascontiguousarray(seg_end, dtype=np.int64)
        self.seg_strand = np.ascontiguousarray(seg_strand, dtype=np.int8)
        self.rec_lit_off = np.ascontiguousarray(rec_lit_off, dtype=np.int64)
        self.rec_pre_len = np.ascontiguousarray(rec_pre_len, dtype=np.int32)
        self.rec_suf_len = np.ascontiguousarray(rec_suf_len, dtype=np.int32)
        self.lit = np.ascontiguousarray(lit, dtype=np.uint8)
        self.rec_phase = None if rec_phase is None else np.ascontiguousarray(rec_phase, dtype=np.int8)

    @property
    def n_rec(self):
        return self.rec_seg_off.size - 1

    @property
    def n_seg(self):
        return self.seg_contig.size

    def slice(self, r0, r1):
        """Records [r0, r1) as an independent table (literal buffer shared, offsets kept)."""
        s0, s1 = int(self.rec_seg_off[r0]), int(self.rec_seg_off[r1])
        return RecordTable(self.rec_seg_off[r0:r1 + 1] - s0, self.seg_contig[s0:s1], self.seg_start[s0:s1],
                           self.seg_end[s0:s1], self.seg_strand[s0:s1], self.rec_lit_off[r0:r1],
                           self.rec_pre_len[r0:r1], self.rec_suf_len[r0:r1], self.lit,
                           None if self.rec_phase is None else self.rec_phase[r0:r1])

    def approx_bytes_per_record(self):
        """Upper estimate of each record's nucleotide text size (before clamping), for shard balancing."""
        seg_len = np.maximum(self.seg_end - self.seg_start + 1, 0)
        csum = np.concatenate(([0], np.cumsum(seg_len)))
        pay = csum[self.rec_seg_off[1:]] - csum[self.rec_seg_off[:-1]]
        return pay + self.rec_pre_len + self.rec_suf_len
