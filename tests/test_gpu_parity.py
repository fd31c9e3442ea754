"""Parity of the CUDA path (through the C ABI / the Python API above it) against the oracle and the
reference-generated goldens.  Everything here is bit-exact (integer / byte work)."""
import contextlib
import io
import os

import numpy as np
import pytest

import coracle
import magot_oracle as mo
from cksum import cksum

pytestmark = pytest.mark.gpu


def _ck(text):
    c, n = cksum(text)
    return {"cksum": c, "bytes": n}


def _stdout_of(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a, **k)
    return buf.getvalue()


@pytest.fixture(scope="module")
def mg():
    import magot_b200
    return magot_b200


# ---- Sequence ops against the reference-generated known answers ---------------------------------------

def test_kat_reverse_compliment(mg, kat):
    for s, r in kat["reverse_compliment"]:
        assert str(mg.Sequence(s).reverse_compliment()) == r


def test_kat_translate_all_frames(mg, kat):
    for s, frame, strand, trimx, r in kat["translate"]:
        if r == "!IndexError":
            continue
        assert mg.Sequence(s).translate(frame=frame, strand=strand, trimX=trimx) == r, (s[:30], frame, strand, trimx)


def test_kat_translate_custom_library(mg, kat):
    """Sequence.translate(library=...) against the reference's own answers (tests/golden/make_golden.py:library_vectors)."""
    libs = kat["translate_library"]["libraries"]
    n = 0
    for lname, s, frame, strand, trimx, r in kat["translate_library"]["vectors"]:
        if r == "!IndexError":
            continue
        assert mg.Sequence(s).translate(library=libs[lname], frame=frame, strand=strand, trimX=trimx) == r, (lname, s[:30], frame, strand)
        n += 1
    assert n > 500
    for bad in ({"ATN": "Q"}, {"ATG": "Met"}):
        with pytest.raises(NotImplementedError):
            mg.Sequence("ATGATG").translate(library=bad)


def test_kat_get_orfs(mg, kat):
    for s, mode, r in kat["get_orfs"][:90]:
        if mode == "longest":
            if r != "!IndexError":
                assert mg.Sequence(s).get_orfs(longest=True) == r
        else:
            assert mg.Sequence(s).get_orfs(from_atg=mode) == r


def test_translate_batch_large(mg):
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    seqs = [alpha[rng.integers(0, 10, size=int(n))].tobytes() for n in rng.integers(0, 4000, size=300)]
    for frame in (0, 1, 2):
        for strand in "+-":
            got = mg.genome._translate_many(seqs, frame=frame, strand=strand)
            for s, g in zip(seqs, got):
                w = coracle.translate(s, frame, strand == '-', True)
                assert g == (None if w is None else w.decode())


# ---- config 1 / config 2: whole files against the reference's cksums ----------------------------------------

def test_suite_goldens_via_entry_points(mg, ref_data, manifest):
    from magot_b200 import genome_tools as gt
    out = _stdout_of(gt.main, ["genome_tools.py", "exclude_from_fasta", os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta"), "NW_011924881.1"])
    assert _ck(out) == manifest["suite:exclude_from_fasta"]
    out = _stdout_of(gt.main, ["genome_tools.py", "cds2pep", os.path.join(ref_data, "CDSannotations.cds")])
    assert _ck(out) == manifest["suite:cds2pep"]
    fa, gtf = os.path.join(ref_data, "C14.fasta"), os.path.join(ref_data, "StandardGTF.gtf")
    out = _stdout_of(gt.main, ["genome_tools.py", "gff2fasta", fa, gtf])
    assert _ck(out) == manifest["suite:gff2fasta_C14_StandardGTF"]
    with open(os.path.join(ref_data, "CDSannotations.cds"), encoding="latin-1", newline="\n") as fh:
        assert out == fh.read()
    out = _stdout_of(gt.main, ["genome_tools.py", "gff2fasta", fa, gtf, "seq_type=protein"])
    assert _ck(out) == manifest["suite:gff2fasta_C14_StandardGTF_protein"]


@pytest.mark.parametrize("gff", ["transcriptlessGTF.gtf", "minimalGFF3.gff"])
def test_c14_annotation_formats(mg, ref_data, manifest, gff):
    from magot_b200 import genome_tools as gt
    fa = os.path.join(ref_data, "C14.fasta")
    assert _ck(_stdout_of(gt.gff2fasta, fa, os.path.join(ref_data, gff))) == manifest["c14:%s" % gff]
    assert _ck(_stdout_of(gt.gff2fasta, fa, os.path.join(ref_data, gff), seq_type="protein")) == manifest["c14:%s:protein" % gff]


def test_c14_genomic_longest_transcript(mg, ref_data, manifest):
    from magot_b200 import genome_tools as gt
    fa, gtf = os.path.join(ref_data, "C14.fasta"), os.path.join(ref_data, "StandardGTF.gtf")
    assert _ck(_stdout_of(gt.gff2fasta, fa, gtf, genomic="True")) == manifest["c14:StandardGTF.gtf:genomic"]
    assert _ck(_stdout_of(gt.gff2fasta, fa, gtf, longest="True")) == manifest["c14:StandardGTF.gtf:longest"]
    g = mg.Genome(fa)
    g.read_gff(gtf)
    assert _ck(g.annotations.get_fasta('transcript') + "\n") == manifest["c14:StandardGTF.gtf:get_fasta_transcript"]


def test_aligner_output_entry_points(mg, ref_data, manifest):
    """blast_csv2fasta / exonerate2fasta (genome_tools.py:265-280) on synthetic aligner outputs against the O.biroi contigs:
    whole stdout against the reference's (cksums from the shimmed reference, tests/golden/make_golden.py)."""
    from magot_b200 import genome_tools as gt
    fa = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    inputs = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aligner_inputs")
    assert _ck(_stdout_of(gt.blast_csv2fasta, fa, os.path.join(inputs, "obiroi_blast.csv"))) == manifest["obiroi:blast_csv2fasta"]
    assert _ck(_stdout_of(gt.exonerate2fasta, fa, os.path.join(inputs, "obiroi_exonerate.txt"))) == manifest["obiroi:exonerate2fasta"]
    g = mg.Genome(fa, os.path.join(inputs, "obiroi_blast.csv"), annotation_format='blast_csv')
    assert _ck(g.annotations.get_fasta('match') + "\n") == manifest["obiroi:blast_csv2fasta"]


def test_mask_from_gff(mg, ref_data, manifest):
    """mask_from_gff (genome_tools.py:394-428) through the K5 interval scatter: soft / hard, with and without upper-casing first,
    on a FASTA with soft-masked, IUPAC and out-of-alphabet bytes, CRLF lines and a repeated header, and a GFF with
    overlapping, duplicate, clamped, negative and empty intervals; whole stdout against the reference's."""
    from magot_b200 import genome_tools as gt
    inputs = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aligner_inputs")
    fa, gff, gff_soft = (os.path.join(inputs, f) for f in ("mask.fasta", "mask.gff", "mask_soft.gff"))
    ob_fa = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    ob_gff = os.path.join(ref_data, "O.biroi_NCBIrefseq_gff3Subset.gff")
    assert _ck(_stdout_of(gt.mask_from_gff, fa, gff_soft)) == manifest["mask:soft"]
    assert _ck(_stdout_of(gt.mask_from_gff, fa, gff_soft, overwrite_softmask="False")) == manifest["mask:soft_keep_case"]
    assert _ck(_stdout_of(gt.mask_from_gff, fa, gff, mask_type="hard")) == manifest["mask:hard"]
    assert _ck(_stdout_of(gt.mask_from_gff, fa, gff, mask_type="hard", overwrite_softmask="F", feature_type="exon")) == \
        manifest["mask:hard_keep_case_exon"]
    assert _ck(_stdout_of(gt.mask_from_gff, ob_fa, ob_gff)) == manifest["mask:obiroi_soft"]
    assert _ck(_stdout_of(gt.mask_from_gff, ob_fa, ob_gff, mask_type="hard", feature_type="exon", overwrite_softmask="False")) == \
        manifest["mask:obiroi_hard_exon"]
    # option handling of the reference: a message and None
    assert "Invalid option for 'overwrite_softmask'" in _stdout_of(gt.mask_from_gff, fa, gff, overwrite_softmask="maybe")
    assert "Invalid option for mask_type" in _stdout_of(gt.mask_from_gff, fa, gff, mask_type="medium")
    with pytest.raises(NotImplementedError):            # the reference would lengthen the sequence here
        gt.mask_from_gff(fa, gff_soft, mask_type="hard")


def test_obiroi_whole_api(mg, ref_data, manifest):
    from magot_b200 import genome_tools as gt
    fa = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    gff = os.path.join(ref_data, "O.biroi_NCBIrefseq_gff3Subset.gff")
    out = _stdout_of(gt.gff2fasta, fa, gff)
    assert out == mo.gff2fasta(fa, gff)
    assert _ck(out) == manifest["obiroi:gff2fasta"]
    assert _ck(_stdout_of(gt.gff2fasta, fa, gff, seq_type="protein")) == manifest["obiroi:gff2fasta_protein"]
    assert _ck(_stdout_of(gt.gff2fasta, fa, gff, from_exons="True")) == manifest["obiroi:gff2fasta_from_exons"]
    g = mg.Genome(fa)
    g.read_gff(gff, base_features=['exon', 'match_part', 'similarity', 'region'], features_to_ignore=['CDS'])
    assert _ck(g.annotations.get_fasta('gene') + "\n") == manifest["obiroi:exon_transcripts"]
    g = mg.Genome(fa, gff, annotation_format='gff3')
    a = g.annotations
    assert _ck(a.get_fasta('mRNA') + "\n") == manifest["obiroi:get_fasta_mRNA"]
    assert _ck(a.get_fasta('mRNA', seq_type="protein") + "\n") == manifest["obiroi:get_fasta_mRNA_protein"]
    assert _ck("\n".join(a.mRNA[k].get_fasta(name_from='Name') for k in a.mRNA) + "\n") == manifest["obiroi:mRNA_name_from_Name"]
    assert _ck("\n".join(a.mRNA[k].get_fasta(genomic=True) for k in a.mRNA) + "\n") == manifest["obiroi:mRNA_genomic"]
    assert _ck(g.get_genome_fasta() + "\n") == manifest["obiroi:genome_fasta"]
    assert _ck("\n".join(k + "\t" + a.CDS[k].get_seq() for k in a.CDS) + "\n") == manifest["obiroi:cds_get_seq"]
    # the reference crashes here (a gene without CDS children): same exception types
    with pytest.raises(TypeError):
        a.get_fasta('gene', genomic=True)
    with pytest.raises(ValueError):
        a.get_fasta('gene', longest=True)
    with pytest.raises(AttributeError):
        a.get_fasta('CDS')


def test_coords2fasta_and_scaffold(mg, ref_data):
    from magot_b200 import genome_tools as gt
    fa = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    for (s, e) in [("1", "80"), ("10000", "12780"), ("0", "5"), ("64000", "99999999"), ("-5", "20")]:
        assert _stdout_of(gt.coords2fasta, fa, "NW_011924877.1", s, e) == mo.coords2fasta(fa, "NW_011924877.1", s, e)
    assert _stdout_of(gt.get_seq_from_fasta, fa, "NW_011924876.1") == mo.get_seq_from_fasta(fa, "NW_011924876.1")


def test_extract_upstream_downstream(mg, ref_data):
    from magot_b200 import genome_tools as gt
    fa = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    gff = os.path.join(ref_data, "O.biroi_NCBIrefseq_gff3Subset.gff")
    for stream in ("up", "down"):
        for n in ("50", "30000"):
            assert _stdout_of(gt.extract_upstream_downstream, fa, gff, n, stream) == mo.extract_upstream_downstream(fa, gff, n, stream)


def test_kat_get_seq_clamping(mg, kat):
    for contig, a, b, strand, r in kat["get_seq"]:
        gs = mg.GenomeSequence(">c1\n" + contig + "\n")
        my = mg.Genome(gs)
        aset = mg.AnnotationSet()
        my.annotations = aset
        aset.genome = my
        base = mg.BaseAnnotation("x", "c1", tuple(sorted((a, b))), "CDS", None, strand, {}, aset)
        assert str(base.get_seq()) == r, (a, b, strand)
        gs.close()


# ---- storage: pack -> fetch round trips, arbitrary bytes ------------------------------------------------------

def test_pack_fetch_roundtrip_arbitrary_bytes(mg):
    rng = np.random.default_rng(9)
    alpha = np.frombuffer(b"ACGTacgtNn-RYKMrykmSWBDHV*. xU\xe9\xff", dtype=np.uint8)
    contigs = {}
    for i, n in enumerate([1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 1000, 4097, 70001]):
        p = np.ones(alpha.size)
        p[:10] = 30
        contigs["c%d" % i] = alpha[rng.choice(alpha.size, size=n, p=p / p.sum())].tobytes().decode("latin-1")
    fasta = "".join(">%s\n%s\n" % kv for kv in contigs.items())
    gs = mg.GenomeSequence(fasta)
    assert gs._engine().primary.n_exceptions > 0
    for k, v in contigs.items():
        assert len(gs[k]) == len(v)
        assert str(gs[k]) == v
        for (a, b) in [(0, 1), (3, 40), (len(v) // 2, len(v)), (len(v) - 1, len(v)), (5, 5)]:
            assert gs[k][a:b] == v[a:b]
        raw = gs._engine().primary.fetch(gs.contig_index(k), 0, len(v), True).decode("latin-1")
        assert raw == mo.reverse_compliment(v)
    assert gs["c11"][-5:] == contigs["c11"][-5:] and gs["c11"][7] == contigs["c11"][7]
    gs.close()


@pytest.mark.gpu
def test_pack_fasta_strips_line_ends_on_device(mg):
    """K0f (mg_genome_pack_fasta): raw record bodies with LF / CRLF / stray CR at random places, lengths around the 16-byte
    lane, the 4096-byte tile and the 64 MB trip (carry of < 32 bases between trips), against bytes.translate on the host."""
    from magot_b200 import engine, _lib
    rng = np.random.default_rng(21)
    alpha = np.frombuffer(b"ACGTacgtNn-RYKM *x", dtype=np.uint8)
    raws = []
    for n in [0, 1, 2, 15, 16, 17, 31, 32, 33, 4095, 4096, 4097, 8191, 70001, 300000]:
        a = alpha[rng.integers(0, alpha.size, n)].copy()
        k = rng.integers(0, n + 1)
        if n:
            a[rng.integers(0, n, k // 7)] = 10
            a[rng.integers(0, n, k // 31)] = 13
        raws.append(a)
    raws.append(np.full(5000, 10, dtype=np.uint8))                     # nothing survives
    raws.append(np.frombuffer(b"ACGT\r\n" * 3000 + b"AC", dtype=np.uint8))
    big = alpha[rng.integers(0, 4, (64 << 20) + 4096 + 77)].copy()      # two trips
    big[60::61] = 10
    big[(64 << 20) - 40:(64 << 20) + 40:3] = 10                         # line ends right at the trip boundary
    raws.append(big)
    want = [a.tobytes().translate(None, b"\r\n") for a in raws]
    g = engine.DeviceGenome([len(w) for w in want], device=0)
    for i, a in enumerate(raws):
        g.pack_fasta(i, a)
    g.finalize()
    for i, w in enumerate(want):
        if len(w) <= 400000:
            assert g.fetch(i, 0, len(w)) == w, i
        else:
            for lo in (0, 12345, (64 << 20) - 2_200_000, len(w) - 70000):
                assert g.fetch(i, lo, lo + 70000) == w[lo:lo + 70000], (i, lo)
    # a body whose number of bases differs from the declared contig length is refused
    g2 = engine.DeviceGenome([10], device=0)
    with pytest.raises(_lib.MagotError):
        g2.pack_fasta(0, np.frombuffer(b"ACGT\nACGT\n", dtype=np.uint8))
    with pytest.raises(_lib.MagotError):
        g2.pack_fasta(0, np.frombuffer(b"ACGTACGTACGT\n", dtype=np.uint8))
    g2.close()
    g.close()


# ---- synthetic twins of configs 3/4 against the C oracle --------------------------------------------------------

def _check_fused(g, tbl, text, textp, other=None):
    """K23 (mg_emit_nuc_prot_*: nucleotide + protein text from one pass, genome.py:704-707) and the one-launch form of all
    products (mg_emit_products_*) must give the very bytes of the single-product kernels (which the caller has compared with
    the oracle), after mg_plan_prepare and after mg_plan_prepare_async."""
    from magot_b200 import engine
    for async_prepare in (False, True):
        n, p = engine.run_table_both(g, tbl, async_prepare=async_prepare)
        assert n == text, ("K23 nucleotide text", async_prepare)
        assert p == textp, ("K23 protein text", async_prepare)
    a, bn, bp = engine.run_products(g, other if other is not None else tbl, tbl)
    assert bn == text and bp == textp
    if other is None:
        assert a == text
    return a


def _oracle_products(contigs, tbl):
    lens = np.array([a.size for a in contigs])
    # Python slice semantics for in-range/overrunning coordinates (all synthetic starts are >= 1)
    lo = np.clip(tbl.seg_start - 1, 0, lens[tbl.seg_contig])
    hi = np.clip(tbl.seg_end, 0, lens[tbl.seg_contig])
    nuc, off = coracle.splice([a.tobytes() for a in contigs], tbl.rec_seg_off, tbl.seg_contig, lo, hi, tbl.seg_strand)
    aa, aa_off, aa_len = coracle.splice_translate(nuc, off)
    return nuc, off, aa, aa_off, aa_len


@pytest.mark.parametrize("kind,total,ntx,seed", [("human", 6_000_000, 3000, 4), ("insect", 3_000_000, 2000, 3)])
def test_synthetic_twin_against_c_oracle(mg, kind, total, ntx, seed):
    from magot_b200 import engine, synth
    layout = synth.contig_layout(kind, total, seed)
    contigs = synth.synth_genome_host(layout, seed, n_mean=300)
    # sprinkle bytes outside the packed alphabet
    rng = np.random.default_rng(seed)
    for a in contigs[:5]:
        idx = rng.integers(0, a.size, size=50)
        a[idx] = np.frombuffer(b"RYKMSWBDHVrykm*", dtype=np.uint8)[rng.integers(0, 15, size=50)]
    ann = synth.synth_annotation(layout, ntx, seed)
    g = engine.DeviceGenome([a.size for a in contigs], device=0)
    for i, a in enumerate(contigs):
        g.pack(i, a)
    g.finalize()
    for which in ("cds", "exon"):
        tbl = ann.table(which, framing=False)
        nuc, off, aa, aa_off, aa_len = _oracle_products(contigs, tbl)
        plan = engine.Plan(g, tbl)
        nt, pt = plan.prepare()
        assert nt == nuc.size and pt == aa.size
        got_n, got_a = plan.lengths()
        assert np.array_equal(got_n, np.diff(off)) and np.array_equal(got_a, aa_len)
        assert plan.emit_host(protein=False).tobytes() == nuc.tobytes()
        assert plan.emit_host(protein=True).tobytes() == aa.tobytes()
        plan.close()
        # with FASTA framing: one D2H yields the final file
        tblf = ann.table(which, framing=True)
        text, _ = engine.run_table(g, tblf)
        want = b"".join(b">" + n.encode() + b"\n" + nuc[off[i]:off[i + 1]].tobytes() + b"\n" for i, n in enumerate(ann.names))
        assert text == want
        textp, _ = engine.run_table(g, tblf, protein=True)
        wantp = b"".join(b">" + n.encode() + b"\n" + aa[aa_off[i]:aa_off[i + 1]].tobytes() + b"\n" for i, n in enumerate(ann.names))
        assert textp == wantp
        _check_fused(g, tbl, nuc.tobytes(), aa.tobytes())
        if which == "cds":                                # the step of config 4: exon text + CDS text + protein in one launch
            exon_f = ann.table("exon", framing=True)
            got_exon = _check_fused(g, tblf, want, wantp, other=exon_f)
            assert got_exon == engine.run_table(g, exon_f)[0]
        else:
            _check_fused(g, tblf, want, wantp)
    g.close()


@pytest.mark.parametrize("framing", [False, True])
def test_dense_tiny_segments_overflow_the_tile_staging(mg, framing):
    """Thousands of 1-8 base segments per 32 KB of text: far more pieces / records per tile than K2 and K3 stage in shared
    memory (NUC_CAP, PROT_RCAP, PROT_PCAP), so whole tiles go through the generic global-memory paths -- same bytes."""
    from magot_b200 import engine
    rng = np.random.default_rng(77)
    alpha = np.frombuffer(b"ACGTacgtNnRY*", dtype=np.uint8)
    contigs = [alpha[rng.integers(0, alpha.size, size=n)].copy() for n in (50_000, 777, 120_000)]
    g = engine.DeviceGenome([a.size for a in contigs], device=0)
    for i, a in enumerate(contigs):
        g.pack(i, a)
    g.finalize()
    n_rec = 4000
    n_seg = rng.integers(0, 40, size=n_rec)
    n_seg[::50] = 400                                   # a few records with hundreds of tiny segments
    rec_off = np.concatenate(([0], np.cumsum(n_seg)))
    E = int(rec_off[-1])
    cid = rng.integers(0, 3, size=E).astype(np.int32)
    lens = np.array([a.size for a in contigs])[cid]
    st = rng.integers(1, lens - 10)
    en = st + rng.integers(0, 8, size=E)
    sd = rng.integers(0, 2, size=E).astype(np.int8)
    names = ["r%d" % i for i in range(n_rec)]
    if framing:
        pre = np.array([len(n) + 2 for n in names], dtype=np.int32)
        suf = np.ones(n_rec, dtype=np.int32)
        lit = np.frombuffer("".join(">" + n + "\n\n" for n in names).encode(), dtype=np.uint8)
        lit_off = np.concatenate(([0], np.cumsum(pre + suf)[:-1]))
    else:
        pre = suf = np.zeros(n_rec, dtype=np.int32)
        lit = np.zeros(0, dtype=np.uint8)
        lit_off = np.zeros(n_rec, dtype=np.int64)
    tbl = engine.RecordTable(rec_off, cid, st, en, sd, lit_off, pre, suf, lit)
    nuc, off, aa, aa_off, aa_len = _oracle_products(contigs, tbl)
    text, (got_n, got_a) = engine.run_table(g, tbl, want_lengths=True)
    textp, _ = engine.run_table(g, tbl, protein=True)
    assert np.array_equal(got_n, np.diff(off)) and np.array_equal(got_a, aa_len)
    if framing:
        want = b"".join(b">" + n.encode() + b"\n" + nuc[off[i]:off[i + 1]].tobytes() + b"\n" for i, n in enumerate(names))
        wantp = b"".join(b">" + n.encode() + b"\n" + aa[aa_off[i]:aa_off[i + 1]].tobytes() + b"\n" for i, n in enumerate(names))
    else:
        want, wantp = nuc.tobytes(), aa.tobytes()
    assert text == want
    assert textp == wantp
    _check_fused(g, tbl, want, wantp)
    g.close()


@pytest.mark.parametrize("seed", list(range(12)))
def test_random_tables_differential(mg, seed):
    """Randomised differential test of K1 + K2 + K3 against the C oracle: contigs of any length (incl. 0) over an alphabet with
    soft-masked, IUPAC and out-of-alphabet bytes; records with 0-12 segments whose coordinates run off either contig end, are
    zero or negative (Python slice rules, genome.py:606), both strands, and framing of 0-40 bytes per record."""
    from magot_b200 import engine
    rng = np.random.default_rng(5000 + seed)
    alpha = np.frombuffer(b"ACGTACGTACGTacgtacgtNnRYKM-*xS", dtype=np.uint8)
    n_contig = int(rng.integers(1, 6))
    contigs = [alpha[rng.integers(0, alpha.size, size=int(n))].copy() for n in rng.choice([0, 1, 2, 3, 31, 32, 33, 100, 1000, 5000, 40000], size=n_contig)]
    g = engine.DeviceGenome([a.size for a in contigs], device=0)
    for i, a in enumerate(contigs):
        g.pack(i, a)
    g.finalize()
    n_rec = int(rng.integers(1, 400))
    n_seg = rng.integers(0, 13, size=n_rec)
    rec_off = np.concatenate(([0], np.cumsum(n_seg)))
    E = int(rec_off[-1])
    cid = rng.integers(0, n_contig, size=E).astype(np.int32)
    L = np.array([a.size for a in contigs], dtype=np.int64)[cid]
    st = rng.integers(-20, L + 30)
    en = st + rng.integers(-5, 700, size=E)
    swap = en < st                                       # the GFF reader sorts each pair (genome.py:309-311)
    st, en = np.where(swap, en, st), np.where(swap, st, en)
    sd = rng.integers(0, 2, size=E).astype(np.int8)
    pre = rng.integers(0, 30, size=n_rec).astype(np.int32)
    suf = rng.integers(0, 12, size=n_rec).astype(np.int32)
    if seed % 3 == 0:
        pre[:] = 0
        suf[:] = 0
    lit = rng.integers(33, 127, size=int(pre.sum() + suf.sum()), dtype=np.uint8)
    lit_off = np.concatenate(([0], np.cumsum(pre.astype(np.int64) + suf)[:-1]))
    tbl = engine.RecordTable(rec_off, cid, st, en, sd, lit_off, pre, suf, lit)
    lo = np.empty(E, dtype=np.int64)
    hi = np.empty(E, dtype=np.int64)
    for e in range(E):                                   # Python slice semantics of contig[start-1:end]
        a, b, _ = slice(int(st[e]) - 1, int(en[e])).indices(int(L[e]))
        lo[e], hi[e] = a, max(a, b)
    nuc, off = coracle.splice([a.tobytes() for a in contigs], rec_off, cid, lo, hi, sd)
    aa, aa_off, aa_len = coracle.splice_translate(nuc, off)
    text, (got_n, got_a) = engine.run_table(g, tbl, want_lengths=True)
    textp, _ = engine.run_table(g, tbl, protein=True)
    assert np.array_equal(got_n, np.diff(off)) and np.array_equal(got_a, aa_len)
    want, wantp = [], []
    for r in range(n_rec):
        p0 = int(lit_off[r])
        head, tail = lit[p0:p0 + pre[r]].tobytes(), lit[p0 + pre[r]:p0 + pre[r] + suf[r]].tobytes()
        want.append(head + nuc[off[r]:off[r + 1]].tobytes() + tail)
        wantp.append(head + aa[aa_off[r]:aa_off[r + 1]].tobytes() + tail)
    assert text == b"".join(want)
    assert textp == b"".join(wantp)
    _check_fused(g, tbl, text, textp)
    g.close()


def test_edge_records(mg):
    """Empty records, zero-length segments, 1-base segments, records <= 2 bases (translate -> None)."""
    from magot_b200 import engine
    contig = np.frombuffer(b"ACGTNNacgtTTGACCATGGGTAAACTGATCGATCGTAGCTAGCTAGCTAGCATCGATCGAT" * 3, dtype=np.uint8)
    g = engine.DeviceGenome([contig.size], device=0)
    g.pack(0, contig)
    g.finalize()
    L = contig.size
    recs = [[], [(5, 4, 0)], [(1, 1, 0)], [(1, 2, 1)], [(1, 3, 1)], [(L, L, 1), (1, 1, 0)], [(L - 1, L + 50, 0)], [(L + 5, L + 9, 0)],
            [(1, L, 1)], [(2, 1, 0), (3, 2, 0), (10, 30, 0)], [(1, 3, 0)] * 40, [(7, 7, 1)] * 33, [(1, L, 0)]]
    rec_off, cid, st, en, sd = [0], [], [], [], []
    for r in recs:
        for (a, b, m) in r:
            cid.append(0); st.append(a); en.append(b); sd.append(m)
        rec_off.append(len(cid))
    R = len(recs)
    z = np.zeros(R, dtype=np.int32)
    tbl = engine.RecordTable(rec_off, cid, st, en, sd, np.zeros(R, np.int64), z, z, np.zeros(0, np.uint8))
    text = contig.tobytes().decode()
    want_n, want_p, want_len = [], [], []
    for r in recs:
        s = "".join(mo.reverse_compliment(text[a - 1:b]) if m else text[a - 1:b] for (a, b, m) in r)
        want_n.append(s)
        t = mo.translate(s)
        want_len.append(-1 if t is None else len(t))
        want_p.append(t or "")
    plan = engine.Plan(g, tbl)
    plan.prepare()
    nl, al = plan.lengths()
    assert list(nl) == [len(s) for s in want_n] and list(al) == want_len
    assert plan.emit_host(False).tobytes().decode() == "".join(want_n)
    assert plan.emit_host(True).tobytes().decode() == "".join(want_p)
    plan.close()
    _check_fused(g, tbl, "".join(want_n).encode(), "".join(want_p).encode())
    empty = engine.RecordTable([0], [], [], [], [], [], [], [], np.zeros(0, np.uint8))
    assert engine.run_table(g, empty)[0] == b""
    g.close()


# ---- K4: six-frame translation + ORF scan -------------------------------------------------------------------------

def _sixframe_case(mg, contigs, min_aa):
    from magot_b200 import engine, orfs, _lib
    g = engine.DeviceGenome([len(c) for c in contigs], device=0)
    for i, c in enumerate(contigs):
        g.pack(i, np.frombuffer(c, dtype=np.uint8))
    g.finalize()
    recs, aa = orfs.sixframe(g, 0, len(contigs), min_aa)
    if min_aa >= 96:                                     # the three scan variants (stop index / packed bases / single pass) agree
        for mode in (1, 0):
            _lib.check(_lib.lib.mg_tune(b"six", mode))
            try:
                recs2, aa2 = orfs.sixframe(g, 0, len(contigs), min_aa)
            finally:
                _lib.check(_lib.lib.mg_tune(b"six", 2))
            assert aa2 == aa and np.array_equal(recs2, recs), mode
    g.close()
    want_aa, want_rec = [], []
    off = 0
    for ci, c in enumerate(contigs):
        a, r = coracle.sixframe(c, min_aa)
        for row in r:
            want_rec.append((ci, int(row[0]), int(row[1]), int(row[2]), int(row[3]), off))
            off += int(row[3])
        want_aa.append(a)
    got_rec = [(int(r["contig"]), int(r["frame"]), int(r["minus"]), int(r["start"]), int(r["len"]), int(r["aa_off"])) for r in recs]
    assert got_rec == want_rec
    assert aa == b"".join(want_aa)


def test_sixframe_matches_reference_get_orfs(mg, kat):
    """Sequence.get_orfs (genome.py:824-851) known answers produced by the reference, via the device ORF scan."""
    from magot_b200 import engine, orfs
    for s, mode, r in kat["get_orfs"]:
        if mode is not False or len(s) < 1:
            continue
        g = engine.DeviceGenome([len(s)], device=0)
        g.pack(0, np.frombuffer(s.encode("latin-1"), dtype=np.uint8))
        g.finalize()
        recs, aa = orfs.sixframe(g, 0, 1, 0)
        g.close()
        text = aa.decode("latin-1")
        got = [text[int(x["aa_off"]):int(x["aa_off"]) + int(x["len"])] for x in recs]
        assert got == r, (len(s), s[:40])


@pytest.mark.parametrize("min_aa", [0, 1, 15, 16, 30, 31, 32, 33, 47, 48, 49, 95, 96, 97, 100, 127, 128, 129, 300, 4000])
def test_sixframe_random_contigs(mg, min_aa):
    rng = np.random.default_rng(100 + min_aa)
    alpha = np.frombuffer(b"ACGTacgtNnR", dtype=np.uint8)
    p = np.array([30, 30, 30, 30, 10, 10, 10, 10, 2, 1, 1], dtype=float)
    contigs = []
    for n in [0, 1, 2, 3, 4, 5, 6, 7, 8, 47, 48, 49, 95, 96, 97, 12287, 12288, 12289, 12290, 24575, 24576, 24577, 24578, 30000, 49151, 49152, 49153, 49154, 73727, 73728, 73729, 98305, 147457, 200001]:
        contigs.append(alpha[rng.choice(alpha.size, size=n, p=p / p.sum())].tobytes())
    # stop-poor contig (long ORFs crossing many tiles) and an N run in the middle of a contig
    contigs.append(np.frombuffer(b"ACG", dtype=np.uint8)[rng.integers(0, 3, size=60000)].tobytes())
    c = bytearray(alpha[rng.integers(0, 4, size=50000)].tobytes())
    c[20000:33000] = b"N" * 13000
    contigs.append(bytes(c))
    # a stop-free stretch longer than several tiles inside a long contig (the two-level scan reads back over tile summaries),
    # and contigs that end / begin inside stop-free windows
    c = bytearray(alpha[rng.integers(0, 4, size=400000)].tobytes())
    c[60000:300000] = np.frombuffer(b"ACG", dtype=np.uint8)[rng.integers(0, 3, size=240000)].tobytes()
    contigs.append(bytes(c))
    contigs.append(np.frombuffer(b"ACG", dtype=np.uint8)[rng.integers(0, 3, size=73728 * 2 + 5)].tobytes())
    _sixframe_case(mg, contigs, min_aa)


def test_prepare_async_equals_prepare(mg):
    """mg_plan_prepare_async (no host round trip, grids sized by the caller's capacities, text sizes read on the device)
    produces the same two texts as mg_plan_prepare -- with tight capacities, generous ones, and with segments that the
    Python-slice clamp shortens (capacity >> real size); too small a capacity is reported by mg_plan_totals."""
    import torch
    from magot_b200 import engine, synth, _lib
    layout = synth.contig_layout("human", 3_000_000, 11)
    contigs = synth.synth_genome_host(layout, 11, n_mean=300)
    g = engine.DeviceGenome([a.size for a in contigs], device=0)
    for i, a in enumerate(contigs):
        g.pack(i, a)
    g.finalize()
    ann = synth.synth_annotation(layout, 1500, 11)
    for which, framing in (("cds", True), ("exon", True), ("cds", False)):
        tbl = ann.table(which, framing=framing)
        # make some segments run off their contig: the clamp shortens them, the host-side upper bound does not know
        tbl.seg_end[::97] += 5_000_000
        plan = engine.Plan(g, tbl)
        nuc, prot = plan.prepare()
        want_n = plan.emit_host(protein=False).tobytes()
        want_p = plan.emit_host(protein=True).tobytes()
        plan.close()
        for slack in (0, 1, 100_000):
            plan = engine.Plan(g, tbl)
            cap_n, cap_p = plan.capacities()
            assert cap_n >= nuc and cap_p >= prot
            if slack == 0:
                cap_n, cap_p = nuc, prot                       # exactly tight
            else:
                cap_n, cap_p = cap_n + slack, cap_p + slack
            plan.prepare_async(cap_n, cap_p)
            out_n = torch.zeros((cap_n + 31) // 32 * 32 + 32, dtype=torch.uint8, device="cuda")
            out_p = torch.zeros((cap_p + 31) // 32 * 32 + 32, dtype=torch.uint8, device="cuda")
            plan.emit_device(out_n.data_ptr(), protein=False)
            plan.emit_device(out_p.data_ptr(), protein=True)
            assert plan.totals() == (nuc, prot)
            assert out_n[:nuc].cpu().numpy().tobytes() == want_n
            assert out_p[:prot].cpu().numpy().tobytes() == want_p
            # host variants after an async prepare resolve the sizes themselves
            assert plan.emit_host(protein=True).tobytes() == want_p
            plan.close()
        # MG_PROT_DEFER: the piece pass alone serves the nucleotide text; protein calls are refused until the record pass has run
        plan = engine.Plan(g, tbl)
        cap_n, cap_p = plan.capacities()
        plan.prepare_async(cap_n, cap_p, defer_records=True)
        out_n = torch.zeros((cap_n + 31) // 32 * 32 + 32, dtype=torch.uint8, device="cuda")
        out_p = torch.zeros((cap_p + 31) // 32 * 32 + 32, dtype=torch.uint8, device="cuda")
        plan.emit_device(out_n.data_ptr(), protein=False)
        with pytest.raises(_lib.MagotError):
            plan.emit_device(out_p.data_ptr(), protein=True)
        assert plan.totals() == (nuc, 0)
        assert out_n[:nuc].cpu().numpy().tobytes() == want_n
        plan.close()
        plan = engine.Plan(g, tbl)
        plan.prepare_async(cap_n, cap_p, defer_records=True)
        plan.prepare_prot()
        plan.emit_device(out_p.data_ptr(), protein=True)
        assert plan.totals() == (nuc, prot)
        assert out_p[:prot].cpu().numpy().tobytes() == want_p
        plan.close()
        plan = engine.Plan(g, tbl)
        plan.prepare_async(max(nuc - 40_000, 0), prot)
        out_n = torch.zeros((nuc + 31) // 32 * 32 + 32, dtype=torch.uint8, device="cuda")
        plan.emit_device(out_n.data_ptr(), protein=False)
        with pytest.raises(_lib.MagotError):
            plan.totals()
        torch.cuda.synchronize()
        assert int(out_n[max(nuc - 40_000, 0) + 64:].max().item()) == 0      # nothing written past the capacity (+ one chunk)
        plan.close()
    g.close()


@pytest.mark.parametrize("min_aa", [0, 30])
def test_sixframe_contig_list_any_order(mg, min_aa):
    """mg_sixframe_count_list: contigs in arbitrary (descending, repeated) order -- the carry between tiles must not leak from a
    contig that lies later in the genome into one that lies earlier; output = the per-contig results in list order."""
    from magot_b200 import engine, orfs
    rng = np.random.default_rng(900 + min_aa)
    alpha = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    contigs = [alpha[rng.integers(0, alpha.size, size=n)].tobytes() for n in (30000, 7, 160000, 73729, 0, 50001)]
    g = engine.DeviceGenome([len(c) for c in contigs], device=0)
    for i, c in enumerate(contigs):
        g.pack(i, np.frombuffer(c, dtype=np.uint8))
    g.finalize()
    per = {}
    for ci, c in enumerate(contigs):
        a, r = coracle.sixframe(c, min_aa)
        per[ci] = (a, [(int(x[0]), int(x[1]), int(x[2]), int(x[3])) for x in r])
    for ids in ([5, 3, 2, 0], [2, 2, 1, 4, 0, 5, 3], [0, 1, 2, 3, 4, 5]):
        recs, aa = orfs.sixframe_list(g, ids, min_aa)
        assert aa == b"".join(per[c][0] for c in ids)
        want = [(c,) + row for c in ids for row in per[c][1]]
        got = [(int(r["contig"]), int(r["frame"]), int(r["minus"]), int(r["start"]), int(r["len"])) for r in recs]
        assert got == want
    shards = orfs.lpt_shards([len(c) for c in contigs], 3)
    assert sorted(c for sh in shards for c in sh) == list(range(len(contigs)))
    g.close()


@pytest.mark.parametrize("min_aa", [0, 16])
def test_sixframe_dense_output_second_pass(mg, min_aa, monkeypatch):
    """More kept ORFs than the scan pass's hit list holds: the records come from the second genome pass instead."""
    monkeypatch.setenv("MG_SIX_HIT_CAP", "7")
    rng = np.random.default_rng(300 + min_aa)
    alpha = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    contigs = [alpha[rng.integers(0, alpha.size, size=n)].tobytes() for n in (5, 73729, 160000)]
    _sixframe_case(mg, contigs, min_aa)


def test_dna2orfs_entry_point(mg, tmp_path):
    """dna2orfs (genome_tools.py:145-180) as intended: ORF strings are get_orfs', positions follow its own formula."""
    from magot_b200 import genome_tools as gt
    rng = np.random.default_rng(7)
    seqs = {"s1": "".join(rng.choice(list("ACGT"), size=900)), "s2 extra words": "".join(rng.choice(list("ACGTN"), size=301))}
    fa = tmp_path / "in.fa"
    fa.write_text("".join(">%s\n%s\n" % kv for kv in seqs.items()))
    out = tmp_path / "orfs.fa"
    gt.dna2orfs(str(fa), str(out))
    got = out.read_text()
    import py2dict
    want = []
    for name in py2dict.py2_order(list(seqs)):
        s = seqs[name]
        for frame in (0, 1, 2):
            for strand in "-+":
                t = mo.translate(s, frame=frame, strand=strand)
                if not t:
                    continue
                pos = frame if strand == '+' else len(s) - frame
                for orf in t.split('*'):
                    want.append(">%s-pos:%d\n%s\n" % (name, pos, orf))
                    pos += 3 * (1 + len(orf))
    assert got == "".join(want)


# ---- configs 3 / 4 end to end through the API: synthetic FASTA + GFF3 / GTF text --------------------------------

@pytest.mark.parametrize("flavour", ["gff3", "gtf"])
def test_synthetic_annotation_text_through_api(mg, tmp_path, flavour):
    """Scaled twins of config 3 (GFF3, gene->mRNA->exon+CDS, CDS rows of one mRNA share an ID: de-dup naming) and
    config 4 (GTF with transcript_id/gene_id): FASTA text + annotation text -> Genome/read_gff/get_fasta vs the oracle."""
    from magot_b200 import synth
    from magot_b200 import genome_tools as gt
    kind = "insect" if flavour == "gff3" else "human"
    layout = synth.contig_layout(kind, 1_500_000, 3)
    contigs = synth.synth_genome_host(layout, 3, n_mean=300)
    ann = synth.synth_annotation(layout, 250, 3)
    names = [n for n, _ in layout]
    fa = tmp_path / "g.fa"
    with open(fa, "wb") as fh:
        for n, a in zip(names, contigs):
            fh.write(b">" + n.encode() + b"\n")
            for k in range(0, a.size, 60):
                fh.write(a[k:k + 60].tobytes() + b"\n")
    gff = tmp_path / ("a." + flavour)
    gff.write_text(ann.to_gff3(names) if flavour == "gff3" else ann.to_gtf(names))
    for kw in ({}, {"seq_type": "protein"}):
        assert _stdout_of(gt.gff2fasta, str(fa), str(gff), **kw) == mo.gff2fasta(str(fa), str(gff), **kw)
    # exon-based transcripts through the library call (base_features=['exon'], CDS ignored)
    g = mg.Genome(str(fa))
    g.read_gff(str(gff), base_features=['exon', 'match_part', 'similarity', 'region'], features_to_ignore=['CDS'])
    seqs, _ = mo.read_fasta(str(fa))
    aset = mo.read_gff(str(gff), base_features=['exon', 'match_part', 'similarity', 'region'], features_to_ignore=['CDS'])
    aset.genome = seqs
    assert g.annotations.get_fasta('gene') == mo.annotation_set_get_fasta(aset, 'gene')
    assert g.annotations.get_fasta('gene', longest=True) == mo.annotation_set_get_fasta(aset, 'gene', longest=True)


# ---- K6: position_dic on the device (genome.py:981-1100) ------------------------------------------------------------

def _position_dic_fixture():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "position_dic.json")) as fh:
        return json.load(fh)


def test_position_dic_at_content_and_windows_match_reference(mg, capsys):
    """at_content (K6 k_at_flags on the packed genome) and sliding_window_calculate (device prefix scan + window gather) against
    the reference's own results: window sums / averages as dicts (with the verbose progress lines), merged regions as an
    AnnotationSet (IDs, coordinates, table order), and mRNA-interval counts over the AT flags."""
    fx = _position_dic_fixture()
    gs = mg.GenomeSequence(fx["fasta"])
    aset = mg.read_gff(fx["gff"])
    pd = mg.position_dic(gs)
    assert list(pd) == fx["order"]
    pd.at_content(gs)
    assert {k: "".join("1" if x else "0" for x in v) for k, v in pd.items()} == fx["at_content"]
    assert pd.count_from_annotations(aset, "mRNA") == fx["count"]["mRNA_over_at"]
    for case in fx["windows"]:
        mg.genome.verbose = case["verbose"]
        capsys.readouterr()
        try:
            r = pd.sliding_window_calculate(case["window"], window_jump=case["jump"], operation=case["operation"],
                                            output=case["output"], threshold=case["threshold"], seqs_to_exclude=case["exclude"])
        finally:
            mg.genome.verbose = True
        assert capsys.readouterr().out == case["stdout"], case
        if case["output"] == "dict":
            assert list(r) == case["key_order"]
            if case["operation"] == "average":
                assert {k: [float(x) for x in v] for k, v in r.items()} == case["result"]      # sums * 1.0 / window: exact
            else:
                assert {k: [int(x) for x in v] for k, v in r.items()} == case["result"]
                assert all(isinstance(x, np.integer) for v in r.values() for x in v[:3])
        else:
            assert [[k, list(v.coords), v.seqid] for k, v in r.region.items()] == case["result"], case
    # "coords" output cannot work in the reference (misspelt name) -- same exceptions here
    with pytest.raises(TypeError):
        pd.sliding_window_calculate(50, output="coords")
    with pytest.raises(NameError):
        pd.sliding_window_calculate(50, output="coords", threshold=[0, 10])
    gs.close()


@pytest.mark.parametrize("dtype", [bool, np.uint8, np.int64, np.int32])
def test_window_sums_random_against_numpy(mg, dtype):
    """mg_window_sums over tile boundaries (4096 / 1024 elements per tile) and several (window, jump) pairs against numpy.sum
    per window, as the reference computes it (genome.py:1055)."""
    rng = np.random.default_rng(12)
    for n in (1, 17, 4095, 4096, 4097, 100_003, 1_300_001):
        if dtype is bool:
            a = rng.random(n) < 0.4
        elif dtype is np.uint8:
            a = rng.integers(0, 256, n).astype(np.uint8)
        else:
            a = rng.integers(-1000, 1000, n).astype(dtype)
        for w, j in ((1, 1), (5, 2), (64, 64), (1000, 37), (n, 1), (n + 5, 3)):
            if n > 200_000 and j < 30:
                continue
            nw = len(range(n)[:-w]) // j if n > w else 0
            if nw == 0:
                continue
            got = mg.position_dic._window_sums(a, w, j, nw)
            want = np.array([int(a[k * j:k * j + w].sum()) for k in range(nw)], dtype=np.int64)
            assert np.array_equal(got.astype(np.int64), want), (n, w, j)
            assert got.dtype == (np.uint64 if dtype is np.uint8 else np.int64)


# ---- in-process multi-GPU: the genome replicated on every device, the record table cut into byte-balanced batches ----------

def test_sharded_genome_on_all_devices_matches_goldens(mg, ref_data, manifest, monkeypatch):
    """SURVEY 8e through the Python API: with DEFAULT_DEVICES = every GPU of the box, gff2fasta (nucleotide and protein) still
    prints the reference suite's goldens -- one host thread per GPU, texts joined in record order.  Skipped on a 1-GPU box."""
    from magot_b200 import genome_tools as gt, _lib
    n = _lib.device_count()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    monkeypatch.setattr(mg.genome, "DEFAULT_DEVICES", list(range(n)))
    fa, gtf = os.path.join(ref_data, "C14.fasta"), os.path.join(ref_data, "StandardGTF.gtf")
    assert _ck(_stdout_of(gt.main, ["genome_tools.py", "gff2fasta", fa, gtf])) == manifest["suite:gff2fasta_C14_StandardGTF"]
    assert _ck(_stdout_of(gt.main, ["genome_tools.py", "gff2fasta", fa, gtf, "seq_type=protein"])) == manifest["suite:gff2fasta_C14_StandardGTF_protein"]
    my = mg.Genome(fa)
    assert len(my.genome_sequence._engine().replicas) == n
    my.read_gff(gtf)
    assert _ck(my.annotations.get_fasta('gene', longest=True) + "\n") == manifest["c14:StandardGTF.gtf:longest"]
    my.genome_sequence.close()


# ---- plan flags: MG_PROT_USE_PHASE and trimX=False ------------------------------------------------------------

@pytest.mark.parametrize("trimx", [True, False])
@pytest.mark.parametrize("use_phase", [True, False])
def test_plan_phase_and_trimx_flags(mg, trimx, use_phase):
    """K1/K3 with MG_PROT_USE_PHASE (the first segment's GFF phase, genome.py:317 keeps it, north_star: "takes the phase
    from the GFF") and with trimX off: record r must be Sequence.translate(trimX=...) (genome.py:795-822) of
    spliced[phase:], phase in {0, 1, 2}; values outside that range count as 0.  The first codon is made of N / IUPAC
    bytes in every third record so that the trimX branch is taken with and without a phase offset."""
    from magot_b200 import engine
    rng = np.random.default_rng(77 + 2 * int(trimx) + int(use_phase))
    alpha = np.frombuffer(b"ACGTACGTACGTacgtNnR", dtype=np.uint8)
    contigs = [alpha[rng.integers(0, alpha.size, size=n)].copy() for n in (5000, 30000, 777)]
    n_rec = 600
    n_seg = rng.integers(1, 7, size=n_rec)
    rec_off = np.concatenate(([0], np.cumsum(n_seg)))
    E = int(rec_off[-1])
    cid = rng.integers(0, 3, size=E).astype(np.int32)
    L = np.array([a.size for a in contigs], dtype=np.int64)[cid]
    st = rng.integers(1, L - 1)
    en = np.minimum(st + rng.integers(0, 300, size=E), L)
    short = rng.random(n_rec) < 0.1                      # records of 0-5 bases: translate -> None around the phase offset
    for r in np.nonzero(short)[0]:
        e0 = rec_off[r]
        en[e0] = st[e0] + rng.integers(0, 5)
        en[e0 + 1:rec_off[r + 1]] = st[e0 + 1:rec_off[r + 1]] - 1      # the other segments are empty
    sd = rng.integers(0, 2, size=E).astype(np.int8)
    phase = rng.choice(np.array([0, 1, 2, 0, 1, 2, -1, 5], dtype=np.int8), size=n_rec)
    for r in range(0, n_rec, 3):                         # N inside the first codon after the phase offset
        e0 = rec_off[r]
        if sd[e0] == 0 and en[e0] - st[e0] > 8:
            contigs[cid[e0]][st[e0] - 1 + max(int(phase[r]), 0) % 3 + int(rng.integers(0, 3))] = ord("N")
    g = engine.DeviceGenome([a.size for a in contigs], device=0)
    for i, a in enumerate(contigs):
        g.pack(i, a)
    g.finalize()
    pre = rng.integers(0, 14, size=n_rec).astype(np.int32)
    suf = np.ones(n_rec, dtype=np.int32)
    lit = rng.integers(33, 127, size=int(pre.sum() + suf.sum()), dtype=np.uint8)
    lit_off = np.concatenate(([0], np.cumsum(pre.astype(np.int64) + suf)[:-1]))
    tbl = engine.RecordTable(rec_off, cid, st, en, sd, lit_off, pre, suf, lit, rec_phase=phase)
    nuc, off = coracle.splice([a.tobytes() for a in contigs], rec_off, cid, st - 1, np.maximum(en, st - 1), sd)
    text, (_, got_a) = engine.run_table(g, tbl, protein=True, trimx=trimx, use_phase=use_phase, want_lengths=True)
    want, want_len = [], []
    for r in range(n_rec):
        ph = int(phase[r]) if (use_phase and 0 <= phase[r] <= 2) else 0
        t = coracle.translate(nuc[off[r]:off[r + 1]].tobytes()[ph:], 0, False, trimx)
        want_len.append(-1 if t is None else len(t))
        p0 = int(lit_off[r])
        want.append(lit[p0:p0 + pre[r]].tobytes() + (t or b"") + lit[p0 + pre[r]:p0 + pre[r] + 1].tobytes())
    assert list(got_a) == want_len
    assert text == b"".join(want)
    assert -1 in want_len and any(w[pre[r]:pre[r] + 1] == b"X" for r, w in enumerate(want))
    g.close()


def test_setitem_on_built_genome(mg):
    """Assigning a contig to a GenomeSequence that is already on the device re-packs it on next use (the reference's
    GenomeSequence is a plain dict, genome.py:854): old contigs keep their text, the new one is served from the device."""
    gs = mg.GenomeSequence(">a\nACGTNNacgt\nTTGA\n>b\nGGGCCC\n")
    assert gs["a"][0:6] == "ACGTNN"
    gs["c"] = "TTTTRYAAAA"
    assert gs["a"][2:12] == "GTNNacgtTT" and str(gs["b"]) == "GGGCCC"
    assert str(gs["c"]) == "TTTTRYAAAA" and gs["c"][4:6] == "RY"
    gs["a"] = "CCCC"                                      # replacing an existing contig keeps its place in the dict
    assert str(gs["a"]) == "CCCC" and str(gs["c"]) == "TTTTRYAAAA" and len(gs) == 3
    assert str(mg.Sequence(gs["c"][0:6]).reverse_compliment()) == "nnAAAA"
    gs.close()


# ---- kernel variants (mg_tune): every K1 / K2 variant gives the same bytes ------------------------------------------

@pytest.fixture()
def tune(mg):
    from magot_b200 import _lib

    def set_(key, v):
        _lib.check(_lib.lib.mg_tune(key.encode(), v))
    yield set_
    set_("emit", 0)
    set_("k1", 0)


@pytest.mark.parametrize("emit,k1", [(0, 1), (0, 0), (1, 1), (2, 1), (2, 0), (1, 0)])
def test_kernel_variants_bit_identical(mg, tune, emit, k1):
    """K2 as k_emit_nuc / k_emit_nuc_tma / k_emit_nuc_stream and K1 as k_plan_rec / the piece-parallel launches, against the C
    oracle: a synthetic twin of config 4 with FASTA framing (hundreds of 32 KB tiles, out-of-alphabet bytes), a table of tiny
    segments (staging overflow paths) and randomised tables with negative / overrunning coordinates."""
    from magot_b200 import engine, synth
    tune("emit", emit)
    tune("k1", k1)
    layout = synth.contig_layout("human", 6_000_000, 4)
    contigs = synth.synth_genome_host(layout, 4, n_mean=300)
    rng = np.random.default_rng(11)
    for a in contigs[:6]:
        idx = rng.integers(0, a.size, size=80)
        a[idx] = np.frombuffer(b"RYKMSWBDHVrykm*", dtype=np.uint8)[rng.integers(0, 15, size=80)]
    ann = synth.synth_annotation(layout, 3000, 4)
    g = engine.DeviceGenome([a.size for a in contigs], device=0)
    for i, a in enumerate(contigs):
        g.pack(i, a)
    g.finalize()
    for which in ("cds", "exon"):
        tbl = ann.table(which, framing=False)
        nuc, off, aa, aa_off, aa_len = _oracle_products(contigs, tbl)
        tblf = ann.table(which, framing=True)
        for sync in (True, False):
            plan = engine.Plan(g, tblf)
            if sync:
                plan.prepare()
            else:
                plan.prepare_async()
                plan.totals()
            got_n, got_a = plan.lengths()
            assert np.array_equal(got_n, np.diff(off)) and np.array_equal(got_a, aa_len)
            text = plan.emit_host(protein=False).tobytes()
            textp = plan.emit_host(protein=True).tobytes()
            plan.close()
            want = b"".join(b">" + n.encode() + b"\n" + nuc[off[i]:off[i + 1]].tobytes() + b"\n" for i, n in enumerate(ann.names))
            wantp = b"".join(b">" + n.encode() + b"\n" + aa[aa_off[i]:aa_off[i + 1]].tobytes() + b"\n" for i, n in enumerate(ann.names))
            assert text == want and textp == wantp, (which, sync)
    # thousands of 1-8 base segments per tile, with and without framing
    n_rec = 300
    n_seg = rng.integers(20, 200, size=n_rec)
    n_seg[7] = 700                                        # longer than one thread of k_plan_rec takes: the piece-parallel K1 runs
    if emit == 2:
        n_seg[7] = 150
    rec_off = np.concatenate(([0], np.cumsum(n_seg)))
    E = int(rec_off[-1])
    cid = rng.integers(0, len(contigs), size=E).astype(np.int32)
    Ls = np.array([a.size for a in contigs], dtype=np.int64)[cid]
    st = rng.integers(1, Ls - 10)
    en = st + rng.integers(0, 8, size=E)
    sd = rng.integers(0, 2, size=E).astype(np.int8)
    for framing in (False, True):
        pre = (rng.integers(1, 40, size=n_rec) if framing else np.zeros(n_rec)).astype(np.int32)
        suf = (np.ones(n_rec) if framing else np.zeros(n_rec)).astype(np.int32)
        lit = rng.integers(33, 127, size=int(pre.sum() + suf.sum()), dtype=np.uint8)
        lit_off = np.concatenate(([0], np.cumsum(pre.astype(np.int64) + suf)[:-1]))
        tbl = engine.RecordTable(rec_off, cid, st, en, sd, lit_off, pre, suf, lit)
        nuc, off = coracle.splice([a.tobytes() for a in contigs], rec_off, cid, st - 1, en, sd)
        aa, aa_off, aa_len = coracle.splice_translate(nuc, off)
        text, _ = engine.run_table(g, tbl)
        textp, _ = engine.run_table(g, tbl, protein=True)
        want, wantp = [], []
        for r in range(n_rec):
            p0 = int(lit_off[r])
            head, tail = lit[p0:p0 + pre[r]].tobytes(), lit[p0 + pre[r]:p0 + pre[r] + suf[r]].tobytes()
            want.append(head + nuc[off[r]:off[r + 1]].tobytes() + tail)
            wantp.append(head + aa[aa_off[r]:aa_off[r + 1]].tobytes() + tail)
        assert text == b"".join(want) and textp == b"".join(wantp), framing
    g.close()
    for seed in (0, 1, 2, 3):
        test_random_tables_differential(mg, seed)
    test_edge_records(mg)


# ---- native reader + flattener (csrc/mg_gff.cu) against the object path ------------------------------------------------

@pytest.mark.parametrize("gff,fasta,kw", [
    ("O.biroi_NCBIrefseq_gff3Subset.gff", "O.biroi_refseqGenomeSubset.fasta", {}),
    ("O.biroi_NCBIrefseq_gff3Subset.gff", "O.biroi_refseqGenomeSubset.fasta", {"base_features": ['exon', 'match_part', 'similarity', 'region'], "features_to_ignore": ['CDS']}),
    ("StandardGTF.gtf", "C14.fasta", {}), ("transcriptlessGTF.gtf", "C14.fasta", {}), ("minimalGFF3.gff", "C14.fasta", {})])
def test_native_flattener_equals_object_path(mg, ref_data, gff, fasta, kw):
    """AnnotationSet.get_fasta straight from the native model (no Python object per feature) == the same call after the
    reference's object model has been built (flatten.py walks the objects), nucleotide and protein, every table."""
    g = mg.Genome(os.path.join(ref_data, fasta))
    g.read_gff(os.path.join(ref_data, gff), **kw)
    aset = g.annotations
    assert aset in mg.genome._PENDING
    native = {}
    for feature in ("gene", "transcript", "mRNA"):
        for seq_type in ("nucleotide", "protein"):
            try:
                native[(feature, seq_type)] = aset.get_fasta(feature, seq_type=seq_type)
            except AttributeError:
                native[(feature, seq_type)] = "AttributeError"
                g.read_gff  # noqa: B018
                if aset not in mg.genome._PENDING:      # an unknown table builds the objects: start again from the text
                    g2 = mg.Genome(g.genome_sequence)
                    g2.read_gff(os.path.join(ref_data, gff), **kw)
                    g, aset = g2, g2.annotations
    g3 = mg.Genome(g.genome_sequence)
    g3.read_gff(os.path.join(ref_data, gff), **kw)
    objs = g3.annotations
    assert len(objs.gene) >= 0 and objs not in mg.genome._PENDING          # touching a table builds the objects
    n_checked = 0
    for (feature, seq_type), text in native.items():
        try:
            want = objs.get_fasta(feature, seq_type=seq_type)
        except AttributeError:
            want = "AttributeError"
        assert text == want, (feature, seq_type)
        n_checked += want not in ("", "AttributeError")
    assert n_checked >= 2
    g.genome_sequence.close()


def test_longest_with_childless_gene(mg, ref_data):
    """gff2fasta longest=True over a set that holds a gene without children: the reference's ParentAnnotation.get_fasta returns
    "" for it before the `longest` block (genome.py:683, :730-731), i.e. a blank line, not ValueError."""
    fa = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    with open(os.path.join(ref_data, "O.biroi_NCBIrefseq_gff3Subset.gff"), encoding="latin-1") as fh:
        lines = [ln for ln in fh.read().split("\n") if ln and not ln.startswith("#") and ln.count("\t") == 8]

    def attr(ln, key):
        for x in ln.split("\t")[8].split(";"):
            if x.startswith(key + "="):
                return x[len(key) + 1:]
        return None
    picked = None
    for gene in (ln for ln in lines if ln.split("\t")[2] == "gene"):
        rna = [ln for ln in lines if attr(ln, "Parent") == attr(gene, "ID") and ln.split("\t")[2] == "mRNA"]
        cds = [ln for ln in lines if ln.split("\t")[2] == "CDS" and attr(ln, "Parent") in {attr(r, "ID") for r in rna}]
        if len(rna) >= 2 and cds:
            picked = [gene] + rna + cds
            break
    assert picked is not None
    lonely = picked[0].replace("ID=" + attr(picked[0], "ID"), "ID=lonely_gene")
    text = "\n".join([lonely] + picked) + "\n"
    g = mg.Genome(fa)
    g.read_gff(text)
    out = g.annotations.get_fasta('gene', longest=True)
    assert out.count(">") == 1 and (out.startswith("\n") or out.endswith("\n"))
    want = mo.read_gff(text)
    seqs, _ = mo.read_fasta(fa)
    want.genome = seqs
    ref = "\n".join(mo.parent_get_fasta(want.table('gene')[k], longest=True) for k in want.table('gene'))
    assert out == ref
    g.genome_sequence.close()
