"""Deterministic synthetic workloads of the shapes BASELINE.json names.

  config 3  "insect":  500 Mbp in 2,000 scaffolds, 30k genes x 2 transcripts, GFF3
  config 4  "human":   3.1 Gbp in 24 chromosomes + 170 scaffolds, 200k transcripts / ~60k genes, GTF
  config 5             six-frame over the config-4 genome
plus scaled-down twins of the same generators for tests and CPU baselines.

Genomes: i.i.d. bases with a GC fraction, runs of N, soft-masked (lower-case) runs.  On the GPU box
the text is produced ON THE DEVICE with torch (plumbing) and packed from there; on the host
(tests, CPU baseline samples) the same shapes come from numpy.  Annotations are produced directly
as SoA interval tables in reference emission order (ascending (start,end), reversed for '-'), and,
for the small twins, also as GFF3/GTF text so the whole API path can be exercised.
"""
import numpy as np

from .tables import RecordTable

GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
          133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
          58617616, 64444167, 46709983, 50818468, 156040895, 57227415]


def contig_layout(kind, total_bp, seed):
    """[(name, length)] for a genome of ~total_bp."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "human":
        n_scaf = 170 if total_bp >= 500_000_000 else max(2, int(170 * total_bp / 3.1e9))
        scaf = rng.integers(50_000, 500_000, size=n_scaf).astype(np.int64)
        if total_bp < 500_000_000:
            scaf = np.maximum(2000, (scaf * (total_bp / 3.1e9) * 4).astype(np.int64))
        rest = max(total_bp - int(scaf.sum()), 24 * 1000)
        chrom = (np.array(GRCH38, dtype=np.float64) / sum(GRCH38) * rest).astype(np.int64)
        names = ["chr%d" % (i + 1) for i in range(22)] + ["chrX", "chrY"] + ["scaffold_%d" % i for i in range(n_scaf)]
        lens = list(chrom) + list(scaf)
    else:
        n = 2000 if total_bp >= 100_000_000 else max(4, int(2000 * total_bp / 5e8))
        w = rng.lognormal(0.0, 1.2, size=n)
        lens = np.maximum(2000, (w / w.sum() * total_bp).astype(np.int64))
        names = ["scaffold_%d" % i for i in range(n)]
    return [(nm, int(l)) for nm, l in zip(names, lens)]


def _runs_mask(rng, n, frac, mean_len):
    """Boolean mask of length n with ~frac of positions covered by geometric runs of mean mean_len."""
    if frac <= 0 or n == 0:
        return np.zeros(n, dtype=bool)
    mean_gap = mean_len * (1 - frac) / frac
    k = int(n / (mean_len + mean_gap) * 1.3) + 8
    gaps = rng.geometric(1.0 / max(mean_gap, 1.0), size=k)
    runs = rng.geometric(1.0 / max(mean_len, 1.0), size=k)
    inter = np.empty(2 * k, dtype=np.int64)
    inter[0::2] = gaps
    inter[1::2] = runs
    vals = np.zeros(2 * k, dtype=bool)
    vals[1::2] = True
    m = np.repeat(vals, inter)
    if m.size < n:
        m = np.concatenate((m, np.zeros(n - m.size, dtype=bool)))
    return m[:n]


def synth_contig_host(length, seed, gc=0.41, n_frac=0.05, soft_frac=0.5, n_mean=2000, soft_mean=500):
    """uint8 ASCII array of one contig (numpy, deterministic in (length, seed))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    at = int(round(256 * (1 - gc) / 2))
    lut = np.empty(256, dtype=np.uint8)
    lut[:at] = ord('A')
    lut[at:2 * at] = ord('T')
    half = (256 - 2 * at) // 2
    lut[2 * at:2 * at + half] = ord('C')
    lut[2 * at + half:] = ord('G')
    a = lut[rng.integers(0, 256, size=length, dtype=np.uint8)]
    soft = _runs_mask(rng, length, soft_frac, soft_mean)
    a[soft] |= 0x20
    nm = _runs_mask(rng, length, n_frac, n_mean)
    a[nm] = ord('N')
    return a


def synth_genome_host(layout, seed, **kw):
    return [synth_contig_host(l, seed * 1000003 + i, **kw) for i, (_, l) in enumerate(layout)]


def synth_contig_device(length, seed, device, gc=0.41, n_frac=0.05, soft_frac=0.5, n_mean=20000, soft_mean=500):
    """torch.uint8 CUDA tensor of one contig's ASCII text (plumbing: workload synthesis only).
    N and soft-mask runs are laid on a coarse grid (run length quantised to 64 bp) to stay cheap."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    at = int(round(256 * (1 - gc) / 2))
    half = (256 - 2 * at) // 2
    lut = torch.empty(256, dtype=torch.uint8, device=device)
    lut[:at] = ord('A')
    lut[at:2 * at] = ord('T')
    lut[2 * at:2 * at + half] = ord('C')
    lut[2 * at + half:] = ord('G')
    r = torch.randint(0, 256, (length,), dtype=torch.uint8, device=device, generator=g)
    a = lut[r.long()] if length < (1 << 24) else torch.take(lut, r.to(torch.int64))
    del r
    nblk = (length + 63) // 64
    # soft-mask: blocks of 64 bp switched on in runs (two-state chain approximated by smoothing random blocks)
    sb = torch.rand(nblk, device=device, generator=g)
    run = max(1, soft_mean // 64)
    sb = torch.nn.functional.avg_pool1d(sb.view(1, 1, -1), kernel_size=2 * run + 1, stride=1, padding=run).view(-1)
    thr = torch.quantile(sb[:min(nblk, 1 << 20)], 1.0 - soft_frac).item() if soft_frac > 0 else 2.0
    soft = (sb > thr).repeat_interleave(64)[:length]
    a = torch.where(soft, a | 0x20, a)
    del soft, sb
    if n_frac > 0:
        nb = torch.rand(nblk, device=device, generator=g)
        runn = max(1, n_mean // 64)
        nb = torch.nn.functional.avg_pool1d(nb.view(1, 1, -1), kernel_size=2 * runn + 1, stride=1, padding=runn).view(-1)
        thr = torch.quantile(nb[:min(nblk, 1 << 20)], 1.0 - n_frac).item()
        isn = (nb > thr).repeat_interleave(64)[:length]
        a = torch.where(isn, torch.full_like(a, ord('N')), a)
    return a.contiguous()


class Annotation(object):
    """Synthetic transcripts as flat arrays, already in reference emission order."""

    def __init__(self, names, gene_of, contig, strand, exon_off, exon_start, exon_end, cds_off, cds_start, cds_end):
        self.names = names                  # transcript ids
        self.gene_of = gene_of              # gene index per transcript
        self.contig = contig                # contig index per transcript
        self.strand = strand                # 0 '+', 1 '-'
        self.exon_off, self.exon_start, self.exon_end = exon_off, exon_start, exon_end
        self.cds_off, self.cds_start, self.cds_end = cds_off, cds_start, cds_end

    @property
    def n_tx(self):
        return len(self.names)

    def table(self, which="cds", framing=True, contig_map=None):
        """RecordTable for the CDS or exon product: record = '>name\\n' + spliced + '\\n'."""
        off, st, en = (self.cds_off, self.cds_start, self.cds_end) if which == "cds" else (self.exon_off, self.exon_start, self.exon_end)
        cnt = np.diff(off)
        contig = np.repeat(self.contig, cnt).astype(np.int32)
        if contig_map is not None:
            contig = np.asarray(contig_map, dtype=np.int32)[contig]
        strand = np.repeat(self.strand, cnt).astype(np.int8)
        T = self.n_tx
        if framing:
            pre = [b">" + n.encode() + b"\n" for n in self.names]
            pre_len = np.fromiter((len(p) for p in pre), dtype=np.int32, count=T)
            lit = np.frombuffer(b"".join(p + b"\n" for p in pre), dtype=np.uint8)
            lit_off = np.concatenate(([0], np.cumsum(pre_len.astype(np.int64) + 1)[:-1]))
            suf_len = np.ones(T, dtype=np.int32)
        else:
            pre_len = np.zeros(T, dtype=np.int32)
            suf_len = np.zeros(T, dtype=np.int32)
            lit = np.zeros(0, dtype=np.uint8)
            lit_off = np.zeros(T, dtype=np.int64)
        return RecordTable(off, contig, st, en, strand, lit_off, pre_len, suf_len, lit)

    def spliced_bp(self, which="cds"):
        st, en = (self.cds_start, self.cds_end) if which == "cds" else (self.exon_start, self.exon_end)
        return int((en - st + 1).sum())

    def subset(self, idx):
        idx = np.asarray(idx, dtype=np.int64)

        def take(off, st, en):
            cnt = np.diff(off)[idx]
            new_off = np.concatenate(([0], np.cumsum(cnt)))
            rows = np.repeat(off[:-1][idx], cnt) + (np.arange(int(new_off[-1])) - np.repeat(new_off[:-1], cnt))
            return new_off, st[rows], en[rows]
        eo, es, ee = take(self.exon_off, self.exon_start, self.exon_end)
        co, cs, ce = take(self.cds_off, self.cds_start, self.cds_end)
        return Annotation([self.names[i] for i in idx], self.gene_of[idx], self.contig[idx], self.strand[idx], eo, es, ee, co, cs, ce)

    def to_gtf(self, contig_names):
        """GTF text with exon + CDS rows and transcript_id / gene_id attributes (config-4 flavour)."""
        lines = []
        for t in range(self.n_tx):
            ctg = contig_names[self.contig[t]]
            sd = "-" if self.strand[t] else "+"
            attr = 'transcript_id "%s"; gene_id "g%d";' % (self.names[t], self.gene_of[t])
            for k in range(self.exon_off[t], self.exon_off[t + 1]):
                lines.append("%s\tsynth\texon\t%d\t%d\t.\t%s\t.\t%s" % (ctg, self.exon_start[k], self.exon_end[k], sd, attr))
            acc = 0
            for k in range(self.cds_off[t], self.cds_off[t + 1]):
                lines.append("%s\tsynth\tCDS\t%d\t%d\t.\t%s\t%d\t%s" % (ctg, self.cds_start[k], self.cds_end[k], sd, (3 - acc % 3) % 3, attr))
                acc += self.cds_end[k] - self.cds_start[k] + 1
        return "\n".join(lines) + "\n"

    def to_gff3(self, contig_names):
        """NCBI-like GFF3: gene -> mRNA -> exon + CDS (config-3 flavour); CDS rows of one mRNA share an ID
        to exercise the reference's de-dup naming."""
        lines = ["##gff-version 3"]
        seen_gene = set()
        for t in range(self.n_tx):
            ctg = contig_names[self.contig[t]]
            sd = "-" if self.strand[t] else "+"
            g = int(self.gene_of[t])
            lo = int(self.exon_start[self.exon_off[t]:self.exon_off[t + 1]].min())
            hi = int(self.exon_end[self.exon_off[t]:self.exon_off[t + 1]].max())
            if g not in seen_gene:
                seen_gene.add(g)
                lines.append("%s\tsynth\tgene\t%d\t%d\t.\t%s\t.\tID=gene%d;Name=G%d" % (ctg, lo, hi, sd, g, g))
            lines.append("%s\tsynth\tmRNA\t%d\t%d\t.\t%s\t.\tID=%s;Parent=gene%d" % (ctg, lo, hi, sd, self.names[t], g))
            for k in range(self.exon_off[t], self.exon_off[t + 1]):
                lines.append("%s\tsynth\texon\t%d\t%d\t.\t%s\t.\tID=ex%d;Parent=%s" % (ctg, self.exon_start[k], self.exon_end[k], sd, k, self.names[t]))
            acc = 0
            for k in range(self.cds_off[t], self.cds_off[t + 1]):
                lines.append("%s\tsynth\tCDS\t%d\t%d\t.\t%s\t%d\tID=cds_%s;Parent=%s" % (ctg, self.cds_start[k], self.cds_end[k], sd, (3 - acc % 3) % 3, self.names[t], self.names[t]))
                acc += self.cds_end[k] - self.cds_start[k] + 1
        return "\n".join(lines) + "\n"


def synth_annotation(layout, n_tx, seed, tx_per_gene=3.3, mean_exons=9.5, exon_median=140, intron_median=1000,
                     cds_frac=0.62, ragged_frac=0.05):
    """Transcripts on the given contigs; everything vectorised (2M exons in well under a second)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([l for _, l in layout], dtype=np.int64)
    n_exon = rng.geometric(1.0 / mean_exons, size=n_tx).astype(np.int64)          # >= 1
    n_exon = np.minimum(n_exon, 120)
    E = int(n_exon.sum())
    off = np.concatenate(([0], np.cumsum(n_exon)))
    exon_len = np.maximum(3, rng.lognormal(np.log(exon_median), 0.8, size=E).astype(np.int64))
    intron_len = np.maximum(30, rng.lognormal(np.log(intron_median), 1.0, size=E).astype(np.int64))
    tx_of = np.repeat(np.arange(n_tx), n_exon)
    first = np.zeros(E, dtype=bool)
    first[off[:-1]] = True
    step = exon_len + intron_len
    pos_in = np.cumsum(step) - step                                           # running start incl. other transcripts
    pos_in -= np.repeat(pos_in[off[:-1]], n_exon)                             # relative start of each exon in its tx
    span = pos_in[off[1:] - 1] + exon_len[off[1:] - 1]
    # contig choice proportional to length, re-drawn onto the largest contig when the span does not fit
    p = lens / lens.sum()
    contig = rng.choice(len(lens), size=n_tx, p=p)
    big = int(np.argmax(lens))
    contig = np.where(span + 2 > lens[contig], big, contig)
    span_fit = np.minimum(span, lens[contig] - 2)
    room = np.maximum(lens[contig] - span_fit - 1, 1)
    tx_start = 1 + (rng.random(n_tx) * room).astype(np.int64)
    strand = rng.integers(0, 2, size=n_tx).astype(np.int8)
    ex_start = np.repeat(tx_start, n_exon) + pos_in
    ex_end = ex_start + exon_len - 1
    # transcripts whose span exceeds even the largest contig run off its end: exercises slice clamping
    # ---- CDS = exons clipped to [a, b) in transcript (spliced, '+'-orientation) coordinates
    sp_end = np.cumsum(exon_len)
    sp_start = sp_end - exon_len
    base = np.repeat(sp_start[off[:-1]], n_exon)
    sp_start -= base
    sp_end -= base
    tx_len = sp_end[off[1:] - 1]
    cds_len = np.maximum(3, (tx_len * cds_frac).astype(np.int64))
    cds_len -= cds_len % 3
    cds_len = np.maximum(cds_len, np.minimum(tx_len, 3))
    ragged = rng.random(n_tx) < ragged_frac
    cds_len = np.where(ragged, np.minimum(tx_len, cds_len + rng.integers(1, 3, size=n_tx)), cds_len)
    a = ((tx_len - cds_len) * rng.random(n_tx)).astype(np.int64)
    b = a + cds_len
    A, B = np.repeat(a, n_exon), np.repeat(b, n_exon)
    lo = np.maximum(sp_start, A)
    hi = np.minimum(sp_end, B)
    keep = hi > lo
    cds_start = (ex_start + (lo - sp_start))[keep]
    cds_end = (ex_start + (hi - sp_start) - 1)[keep]
    cds_cnt = np.bincount(tx_of[keep], minlength=n_tx)
    cds_off = np.concatenate(([0], np.cumsum(cds_cnt)))

    # emission order: ascending for '+', descending for '-' (reverse each transcript's block)
    def emit_order(o, cnt, st, en):
        n = int(o[-1])
        idx = np.arange(n)
        t_of = np.repeat(np.arange(n_tx), cnt)
        rev = np.repeat(strand.astype(bool), cnt)
        within = idx - np.repeat(o[:-1], cnt)
        src = np.where(rev, np.repeat(o[:-1] + cnt - 1, cnt) - within, idx)
        return st[src], en[src]
    ex_start, ex_end = emit_order(off, n_exon, ex_start, ex_end)
    cds_start, cds_end = emit_order(cds_off, cds_cnt, cds_start, cds_end)
    n_gene = max(1, int(n_tx / tx_per_gene))
    gene_of = np.sort(rng.integers(0, n_gene, size=n_tx))
    names = ["tx%07d" % i for i in range(n_tx)]
    return Annotation(names, gene_of, contig.astype(np.int64), strand, off, ex_start, ex_end, cds_off, cds_start, cds_end)
