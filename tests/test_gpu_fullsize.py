"""Parity at BASELINE.json's FULL sizes (config 3: 500 Mbp / 60 k transcripts; configs 4 and 5: 3.1 Gbp genome, 200 k
transcripts, whole-genome six-frame).

WHOLE-TEXT parity against the C oracle (oracle/oracle.c) on HOST-KNOWN bytes: the synthetic genome text is copied to the host
by torch BEFORE the library packs it, the oracle splices / translates ALL records of configs 3 and 4 from those bytes, and
every byte of the three products (CDS nucleotides, CDS protein, exon transcripts; with and without FASTA framing) is
compared; config 5's full ORF list of the 170 scaffolds and one chromosome is compared with the oracle's six-frame scan.

Plus size-independent properties:

  * sampled records: K2's text of a record == its framing + the record's segments fetched one by one through the
    independent range-decode path (mg_genome_fetch, strand-aware) and, for those records, K3's protein == the C oracle's
    translation (oracle/oracle.c, Sequence.translate genome.py:795-822) of K2's own text;
  * byte counts: totals == sum of the Python-slice-clamped segment lengths + framing;
  * idempotence: a second prepare + emit gives the same bytes;
  * sharding linearity: the table cut at arbitrary records and emitted piecewise == the whole text (every cut moves all
    tile boundaries);
  * six-frame: contig ranges concatenate (the multi-GPU shard rule), every ORF has >= min_aa residues, holds no '*', and
    sampled ORFs equal the oracle's translation of the fetched bases.
"""
import ctypes
import zlib

import numpy as np
import pytest

import coracle

pytestmark = pytest.mark.gpu

GENOME_BP = 3_100_000_000
N_TX = 200_000
SEED = 4


def _build(kind, genome_bp, n_tx, seed):
    import torch
    from magot_b200 import engine, synth, _lib
    _lib.require_device(0)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    layout = synth.contig_layout(kind, genome_bp, seed)
    g = engine.DeviceGenome([l for _, l in layout], device=0)
    CH = 256 << 20
    host = []                                            # the genome text as torch produced it, never touched by the library
    for ci, (_, L) in enumerate(layout):
        h = np.empty(L, dtype=np.uint8)
        for off in range(0, L, CH):
            n = min(CH, L - off)
            a = synth.synth_contig_device(n, seed * 1000003 + ci * 64 + off // CH, dev)
            h[off:off + n] = a.cpu().numpy()
            g.pack_device(ci, a.data_ptr(), n, offset=off)
            torch.cuda.synchronize()
            del a
        host.append(h)
    g.finalize()
    torch.cuda.empty_cache()
    _HOST[id(g)] = host
    return g, layout, synth.synth_annotation(layout, n_tx, seed)


_HOST = {}


@pytest.fixture(scope="module")
def insect():
    """BASELINE config 3: 500 Mbp in 2 000 scaffolds, 60 k transcripts."""
    g, layout, ann = _build("insect", 500_000_000, 60_000, 3)
    yield g, layout, ann
    g.close()


@pytest.fixture(scope="module")
def big():
    """BASELINE configs 4 and 5: 3.1 Gbp in 24 chromosomes + 170 scaffolds, 200 k transcripts."""
    g, layout, ann = _build("human", GENOME_BP, N_TX, SEED)
    yield g, layout, ann
    g.close()


def _emit(g, table, protein=False):
    from magot_b200 import engine
    plan = engine.Plan(g, table)
    try:
        plan.prepare()
        nuc_len, aa_len = plan.lengths()
        text = plan.emit_host(protein=protein)
        return np.array(text, copy=True), nuc_len, aa_len
    finally:
        plan.close()


def _record_offsets(table, nuc_len):
    """Start of every record in the nucleotide text: prefix + spliced payload + suffix, back to back."""
    sizes = table.rec_pre_len.astype(np.int64) + nuc_len + table.rec_suf_len.astype(np.int64)
    return np.concatenate(([0], np.cumsum(sizes)))


def _check_whole_text(fixture, which):
    """Every byte of the products of ALL records against the C oracle run on the host copy of the genome text."""
    g, layout, ann = fixture
    host = _HOST[id(g)]
    lens = np.array([l for _, l in layout], dtype=np.int64)
    tbl = ann.table(which, framing=False)
    lo = np.clip(tbl.seg_start - 1, 0, lens[tbl.seg_contig])
    hi = np.clip(tbl.seg_end, 0, lens[tbl.seg_contig])
    nuc, off = coracle.splice(host, tbl.rec_seg_off, tbl.seg_contig, lo, hi, tbl.seg_strand)
    assert nuc.size == ann.spliced_bp(which)
    text, nuc_len, aa_len = _emit(g, tbl)
    assert np.array_equal(nuc_len, np.diff(off))
    assert text.size == nuc.size and np.array_equal(text, nuc), "spliced nucleotides (%s, all records) differ from the oracle" % which
    del text
    aa = aa_off = None
    if which == "cds":
        aa, aa_off, want_aa_len = coracle.splice_translate(nuc, off)
        prot, _, _ = _emit(g, tbl, protein=True)
        assert np.array_equal(aa_len, want_aa_len)
        assert prot.size == aa.size and np.array_equal(prot, aa), "protein text (all records) differs from the oracle"
        del prot
    # the same with FASTA framing: '>' + name + newline + payload + newline per record, compared record by record
    tblf = ann.table(which, framing=True)
    for protein in ((False, True) if which == "cds" else (False,)):
        textf, _, _ = _emit(g, tblf, protein=protein)
        pay, poff = (aa, aa_off) if protein else (nuc, off)
        mv, pv = memoryview(textf), memoryview(pay)
        q = 0
        for r, name in enumerate(ann.names):
            head = b">" + name.encode() + b"\n"
            n = int(poff[r + 1] - poff[r])
            assert mv[q:q + len(head)] == head, ("header", r)
            q += len(head)
            assert mv[q:q + n] == pv[int(poff[r]):int(poff[r]) + n], ("payload", r, protein)
            q += n
            assert textf[q] == 10, ("newline", r)
            q += 1
        assert q == textf.size
        del textf


def test_config3_whole_text_against_oracle(insect):
    _check_whole_text(insect, "cds")
    _check_whole_text(insect, "exon")


@pytest.mark.parametrize("which", ["cds", "exon"])
def test_config4_whole_text_against_oracle(big, which):
    _check_whole_text(big, which)


def test_config5_orf_list_against_oracle(big):
    """The complete ORF list (order, frame, strand, start, length, residues) of every scaffold and of the smallest chromosome
    of the 3.1 Gbp genome, min 100 aa, against the C oracle's Sequence.get_orfs restatement on the host copy of the text."""
    from magot_b200 import orfs
    g, layout, _ = big
    host = _HOST[id(g)]
    lens = [l for _, l in layout]
    chrom = int(np.argmin(lens[:24]))
    ids = [chrom] + list(range(24, len(layout)))
    recs, aa = orfs.sixframe_list(g, ids, 100)
    want_aa, want = [], []
    o = 0
    for c in ids:
        a, rr = coracle.sixframe(host[c], 100)
        for row in rr:
            want.append((c, int(row[0]), int(row[1]), int(row[2]), int(row[3]), o))
            o += int(row[3])
        want_aa.append(a)
    got = list(zip(recs["contig"].tolist(), recs["frame"].tolist(), recs["minus"].tolist(), recs["start"].tolist(),
                   recs["len"].tolist(), recs["aa_off"].tolist()))
    assert len(got) == len(want) and len(want) > 20_000
    assert got == want
    assert aa == b"".join(want_aa)


def test_config3_full_size_properties(insect):
    _check_splice_properties(insect, "cds")


@pytest.mark.parametrize("which", ["cds", "exon"])
def test_config4_full_size_properties(big, which):
    _check_splice_properties(big, which)


def _check_splice_properties(fixture, which):
    g, layout, ann = fixture
    table = ann.table(which)
    lens = np.array([l for _, l in layout], dtype=np.int64)
    text, nuc_len, aa_len = _emit(g, table)
    # byte counts from the Python-slice clamp (genome.py:606) done on the host
    lo = np.clip(table.seg_start - 1, 0, lens[table.seg_contig])
    hi = np.clip(table.seg_end, 0, lens[table.seg_contig])
    seg_len = np.maximum(hi - lo, 0)
    csum = np.concatenate(([0], np.cumsum(seg_len)))
    want_len = csum[table.rec_seg_off[1:]] - csum[table.rec_seg_off[:-1]]
    assert np.array_equal(nuc_len, want_len)
    off = _record_offsets(table, nuc_len)
    assert off[-1] == text.size
    assert ann.spliced_bp(which) == int(want_len.sum())
    # idempotence
    again, _, _ = _emit(g, table)
    assert zlib.crc32(again.tobytes()) == zlib.crc32(text.tobytes())
    del again
    # sampled records against the independent fetch path; first, last and 400 random ones
    rng = np.random.default_rng(17)
    sample = np.unique(np.concatenate(([0, table.n_rec - 1], rng.integers(0, table.n_rec, size=400))))
    prot = None
    if which == "cds":
        prot, _, _ = _emit(g, table, protein=True)
        p_sizes = table.rec_pre_len.astype(np.int64) + np.maximum(aa_len, 0) + table.rec_suf_len.astype(np.int64)
        p_off = np.concatenate(([0], np.cumsum(p_sizes)))
        assert p_off[-1] == prot.size
    for r in sample:
        r = int(r)
        pre, suf = int(table.rec_pre_len[r]), int(table.rec_suf_len[r])
        l0 = int(table.rec_lit_off[r])
        parts = [table.lit[l0:l0 + pre].tobytes()]
        for e in range(int(table.rec_seg_off[r]), int(table.rec_seg_off[r + 1])):
            if seg_len[e] > 0:
                parts.append(g.fetch(int(table.seg_contig[e]), int(lo[e]), int(hi[e]), bool(table.seg_strand[e])))
        parts.append(table.lit[l0 + pre:l0 + pre + suf].tobytes())
        want = b"".join(parts)
        got = text[off[r]:off[r + 1]].tobytes()
        assert got == want, ("record", r)
        if prot is not None:
            payload = got[pre:len(got) - suf]
            w = coracle.translate(payload, 0, False, True)
            gp = prot[p_off[r]:p_off[r + 1]].tobytes()
            if w is None:
                assert aa_len[r] == -1 and gp == want[:pre] + want[len(want) - suf:]
            else:
                assert gp == want[:pre] + w + want[len(want) - suf:], ("protein", r)
    # sharding linearity: cut at arbitrary records, emit the pieces, compare checksums of the concatenation
    cuts = [0] + sorted(int(x) for x in rng.integers(1, table.n_rec - 1, size=3)) + [table.n_rec]
    crc = 0
    total = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        piece, _, _ = _emit(g, table.slice(a, b))
        crc = zlib.crc32(piece.tobytes(), crc)
        total += piece.size
    assert total == text.size and crc == zlib.crc32(text.tobytes())


def test_config5_full_size_properties(big):
    from magot_b200 import orfs
    g, layout, _ = big
    nc = len(layout)
    min_aa = 100
    recs, aa = orfs.sixframe(g, 0, nc, min_aa)
    assert recs.size > 500_000
    assert int(recs["len"].min()) >= min_aa
    assert np.array_equal(recs["aa_off"], np.concatenate(([0], np.cumsum(recs["len"])[:-1])))
    assert int(recs["len"].sum()) == len(aa)
    buf = np.frombuffer(aa, dtype=np.uint8)
    assert not (buf == ord("*")).any()
    # contig ranges concatenate: what each GPU of a multi-GPU run produces is a slice of the whole
    k = nc // 3
    r1, a1 = orfs.sixframe(g, 0, k, min_aa)
    r2, a2 = orfs.sixframe(g, k, nc, min_aa)
    assert a1 + a2 == aa
    assert r1.size + r2.size == recs.size
    for f in ("contig", "frame", "minus", "start", "len"):
        assert np.array_equal(np.concatenate((r1[f], r2[f])), recs[f]), f
    # reference order inside a contig: frame-major, '-' before '+' (genome.py:829-830), start ascending inside a stream
    c0 = recs[recs["contig"] == 0]
    key = c0["frame"].astype(np.int64) * 2 + (1 - c0["minus"].astype(np.int64))
    assert np.all(np.diff(key) >= 0)
    same = np.diff(key) == 0
    assert np.all(np.diff(c0["start"])[same] > 0)
    # sampled ORFs: residues == the oracle's translation of the fetched bases
    rng = np.random.default_rng(5)
    lens = [l for _, l in layout]
    for i in rng.integers(0, recs.size, size=300):
        r = recs[int(i)]
        c, L = int(r["contig"]), lens[int(r["contig"])]
        f, minus, st, ln = int(r["frame"]), bool(r["minus"]), int(r["start"]), int(r["len"])
        # the translation of stream (f, strand) starts at oriented offset cs in {0 or 3 (trimmed X), 2, 4}; residue st is the
        # codon at oriented offset cs + 3*st, so cs is recovered from the stop that precedes the ORF: fetch a window wide
        # enough for every cs and pick the one whose translation has no stop
        got = aa[int(r["aa_off"]):int(r["aa_off"]) + ln]
        ok = False
        for cs in ((0, 3) if f == 0 else ((2,) if f == 1 else (4,))):
            q0 = cs + 3 * st
            if q0 + 3 * ln > L:
                continue
            if minus:
                seq = g.fetch(c, L - q0 - 3 * ln, L - q0, True)
            else:
                seq = g.fetch(c, q0, q0 + 3 * ln, False)
            if coracle.translate(seq, 0, False, False) == got:
                ok = True
                break
        assert ok, (int(i), c, f, minus, st, ln)


def test_config3_fasta_text_ingest_on_device(insect):
    """K0f at config-3 size: the 500 Mbp genome written out as FASTA text (60-column lines, CRLF on every 7th scaffold) and read
    back through GenomeSequence -- headers parsed on the host, bodies stripped and packed on the device -- must decode to the
    same bases as the genome it was written from (whole-contig CRC for the largest scaffolds, sampled ranges for all)."""
    from magot_b200 import genome as mg
    g, layout, _ = insect
    parts = []
    for ci, (name, L) in enumerate(layout):
        seq = np.frombuffer(g.fetch(ci, 0, L), dtype=np.uint8)
        eol = b"\r\n" if ci % 7 == 0 else b"\n"
        full = L // 60
        rows = np.empty((full, 60 + len(eol)), dtype=np.uint8)
        rows[:, :60] = seq[:full * 60].reshape(full, 60)
        rows[:, 60:] = np.frombuffer(eol, dtype=np.uint8)
        parts.append(b">" + name.encode() + b" synthetic scaffold\n" + rows.tobytes() + seq[full * 60:].tobytes() + eol)
    text = b"".join(parts)
    del parts
    gs = mg.GenomeSequence(text)
    try:
        assert len(gs) == len(layout)
        eng = gs._engine().primary
        rng = np.random.default_rng(5)
        order = np.argsort([-l for _, l in layout])
        for ci in range(len(layout)):
            name, L = layout[ci]
            key = name + " synthetic scaffold"
            assert len(gs[key]) == L
            cj = gs.contig_index(key)
            if ci in order[:8]:
                assert zlib.crc32(eng.fetch(cj, 0, L)) == zlib.crc32(g.fetch(ci, 0, L)), name
            else:
                lo = int(rng.integers(0, max(L - 500, 1)))
                hi = min(L, lo + 500)
                assert eng.fetch(cj, lo, hi) == g.fetch(ci, lo, hi), (name, lo)
                assert eng.fetch(cj, max(L - 70, 0), L, True) == g.fetch(ci, max(L - 70, 0), L, True), name
    finally:
        gs.close()
