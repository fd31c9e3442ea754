"""TEST INFRASTRUCTURE -- surrogate for the reference's missing test_data/C14.fasta.

`C14.fasta` is referenced by test_data/test_suite.py:9,12,13 but absent from the checkout
(/root/reference/.MISSING_LARGE_BLOBS).  The expected stdout of
`gff2fasta C14.fasta StandardGTF.gtf` IS shipped (test_data/CDSannotations.cds, cksum
2836090577 690750), so every base under a CDS of StandardGTF.gtf is known: put each golden
record back onto Chromosome14 at its GTF coordinates (un-reverse-complementing '-' segments)
and fill everything else with 'N'.  Any implementation run on this surrogate with
StandardGTF.gtf / minimalGFF3.gff / transcriptlessGTF.gtf must reproduce the golden file.
"""
import gzip

_RC = {'a': 't', 't': 'a', 'g': 'c', 'c': 'g', 'A': 'T', 'T': 'A', 'G': 'C', 'C': 'G', 'n': 'n', 'N': 'N'}
LENGTH = 8589052


def _open(path):
    return gzip.open(path, "rt", encoding="latin-1", newline="\n") if str(path).endswith(".gz") \
        else open(path, encoding="latin-1", newline="\n")


def build(gtf_path, cds_path, length=LENGTH):
    """Return the surrogate FASTA text ('>Chromosome14\\n<one line>\\n')."""
    golden = {}
    name = None
    with _open(cds_path) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                name = line[1:]
                golden[name] = ""
            elif name is not None:
                golden[name] += line
    segs = {}
    with _open(gtf_path) as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if len(f) != 9 or f[2] != "CDS":
                continue
            tid = f[8].split('transcript_id "')[1].split('"')[0]
            segs.setdefault(tid, []).append((int(f[3]), int(f[4]), f[6]))
    chrom = bytearray(b"N" * length)
    for tid, lst in segs.items():
        lst = sorted(set(lst))
        strand = lst[-1][2]
        if strand == '-':
            lst = lst[::-1]
        seq = golden[tid]
        pos = 0
        for (s, e, st) in lst:
            n = e - s + 1
            piece = seq[pos:pos + n]
            pos += n
            if st == '-':
                piece = "".join(_RC[c] for c in reversed(piece))
            old = chrom[s - 1:e]
            new = piece.encode("latin-1")
            assert len(new) == n, (tid, s, e)
            for o, w in zip(old, new):
                assert o == 78 or o == w, "conflicting bases in surrogate"
            chrom[s - 1:e] = new
        assert pos == len(seq), tid
    return ">Chromosome14\n" + chrom.decode("latin-1") + "\n"
