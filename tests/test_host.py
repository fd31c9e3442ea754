"""Host-side logic of the product (no GPU): the C-ABI library loads and exports every symbol the header
declares, fails loudly without a device, and the Python host code (GFF reader, FASTA parser, py2 order,
flattener, sharding) agrees with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import magot_oracle as mo
import py2dict as oracle_py2
from magot_b200 import _lib, engine, genome, py2dict, synth
from magot_b200.flatten import Flattener

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = _lib.device_count() > 0


def test_abi_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "magot_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    declared = set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.mg_version() == 100


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device behaviour")
def test_fails_loudly_without_device():
    n = ctypes.c_int(-1)
    rc = _lib.lib.mg_device_count(ctypes.byref(n))
    assert rc != 0 or n.value == 0
    with pytest.raises(_lib.MagotError):
        engine.DeviceGenome([100], device=0)
    with pytest.raises(_lib.MagotError):
        genome.Sequence("ACGT").reverse_compliment()
    with pytest.raises(_lib.MagotError):
        genome.Sequence("ACGTACGT").translate()
    h = ctypes.c_void_p()
    lens = np.array([10], dtype=np.int64)
    assert _lib.lib.mg_genome_create(0, 1, ctypes.c_void_p(lens.ctypes.data), ctypes.byref(h)) != 0
    assert "no CPU fallback" in _lib.last_error() or "cuda" in _lib.last_error().lower()


def test_py2_order_matches_oracle_emulator():
    rng = np.random.default_rng(1)
    keys = ["g%d.t%d" % (rng.integers(0, 10 ** 6), i) for i in range(5000)] + ["", "a", "x" * 300, "é".encode("utf-8").decode("latin-1")]
    assert py2dict.py2_order(keys) == oracle_py2.py2_order(keys)
    assert py2dict.py2_order_after_deepcopy(keys[:700]) == oracle_py2.py2_order_after_deepcopy(keys[:700])
    hs = py2dict.string_hashes(keys[:50] + ["", "a", "abc"])
    assert hs[-3:] == [0, 12416037344, 1453079729188098211]
    # the native replay (mg_py2_order, used for >= 64 str keys) and the pure-Python one are the same algorithm
    assert py2dict._native_perm(keys, 1) is not None
    assert py2dict._py2_order_python(keys) == oracle_py2.py2_order(keys)
    assert py2dict.py2_order(keys[:10]) == oracle_py2.py2_order(keys[:10])          # short lists take the Python path
    assert py2dict.py2_order([b"ab", "cd"] + keys[:80]) == py2dict._py2_order_python([b"ab", "cd"] + keys[:80])   # non-str key: Python path


def test_py2_order_large_growth_rule():
    # above 50000 entries CPython 2.7 doubles instead of quadrupling; both emulators must agree there
    keys = ["tx%07d" % i for i in range(70000)]
    assert py2dict.py2_order(keys) == oracle_py2.py2_order(keys)
    assert py2dict.py2_order_after_deepcopy(keys) == oracle_py2.py2_order_after_deepcopy(keys)


def _model_signature_product(aset):
    sig = {}
    for name in aset._dict_names():
        for k, o in aset.__dict__[name].items():
            sig[(name, k)] = (o.seqid, o.strand, o.parent, getattr(o, "coords", None), tuple(getattr(o, "child_list", ())),
                              getattr(o, "phase", None))
    order = {name: list(aset.__dict__[name]) for name in aset._dict_names()}
    return sig, order


def _model_signature_oracle(aset):
    sig = {}
    for name in sorted(aset.tables):
        for k, o in aset.tables[name].items():
            sig[(name, k)] = (o.seqid, o.strand, o.parent, getattr(o, "coords", None), tuple(getattr(o, "child_list", ())),
                              getattr(o, "phase", None))
    order = {name: list(aset.tables[name]) for name in sorted(aset.tables)}
    return sig, order


@pytest.mark.parametrize("gff,kw", [
    ("O.biroi_NCBIrefseq_gff3Subset.gff", {}),
    ("O.biroi_NCBIrefseq_gff3Subset.gff", {"base_features": ['exon', 'match_part', 'similarity', 'region'], "features_to_ignore": ['CDS']}),
    ("O.biroi_NCBIrefseq_gff3Subset.gff", {"features_to_ignore": "CDS", "features_to_replace": [('exon', 'CDS')]}),
    ("StandardGTF.gtf", {}), ("transcriptlessGTF.gtf", {}), ("minimalGFF3.gff", {})])
def test_read_gff_matches_oracle(ref_data, gff, kw):
    path = os.path.join(ref_data, gff)
    a = genome.read_gff(path, **kw)
    b = mo.read_gff(path, **kw)
    sa, oa = _model_signature_product(a)
    sb, ob = _model_signature_oracle(b)
    assert sa == sb
    assert oa == ob          # dict iteration order == py2 order after deepcopy


def test_read_gff_dedup_names(ref_data):
    a = genome.read_gff(os.path.join(ref_data, "transcriptlessGTF.gtf"))
    assert {"g3360-CDS", "g3360-CDS2", "g3360-CDS-3"} <= set(a.CDS)
    assert a["g3360"].child_list[:3] == ["g3360-CDS", "g3360-CDS2", "g3360-CDS-3"]
    assert a.transcript == {}


def _bodies(data, names, spans):
    """What K0f leaves of every record body: CR / LF removed, everything else kept."""
    return {n: data[spans[n][0]:spans[n][1]].translate(None, b"\r\n").decode("latin-1") for n in names}


def test_scan_fasta_matches_oracle(ref_data, tmp_path):
    """Header discovery, record rules and lengths of the host FASTA scan (the bodies themselves are stripped on the device)."""
    for text in [">a desc here\nACGT\nacgtNN\n>b\n\n>c\r\nAC GT\r\nRYK-*\n>a desc here\nTTTT\n", "ACGT\n>x\nAC\n", ">only\n", "",
                 ">t1 x\nAAAA\n>t2\tz\nCC\rCC\n", ">e\n\r\n\n>f\nA", "\n\n>g\nAC\n>h"]:
        data = text.encode("latin-1")
        names, spans = genome._scan_fasta(data, False)
        seqs, order = mo.read_fasta(text) if text else ({}, [])
        got = _bodies(data, names, spans)
        assert got == seqs
        assert all(spans[n][2] == len(seqs[n]) for n in names)
        assert py2dict.py2_order(names) == order
    names, spans = genome._scan_fasta(b">t1 x\nAAAA\n>t2\tz\nCC\n", True)
    assert names == ["t1", "t2"]
    path = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    with open(path, "rb") as fh:
        data = fh.read()
    names, spans = genome._scan_fasta(data, False)
    seqs, order = mo.read_fasta(path)
    assert py2dict.py2_order(names) == order
    assert _bodies(data, names, spans) == {n: seqs[n] for n in names}
    assert all(spans[n][2] == len(seqs[n]) for n in names)


class _FakeGS(object):
    """Stands in for GenomeSequence so the flattener can be checked without a device."""
    def __init__(self, names):
        self.idx = {n: i for i, n in enumerate(names)}

    def contig_index(self, seqid):
        return self.idx[seqid]


class _FakeGenome(object):
    def __init__(self, gs):
        self.genome_sequence = gs


def test_flattener_emission_order_matches_oracle(ref_data):
    """Intervals the flattener hands to the device == intervals the oracle slices, record by record."""
    fa = os.path.join(ref_data, "O.biroi_refseqGenomeSubset.fasta")
    gff = os.path.join(ref_data, "O.biroi_NCBIrefseq_gff3Subset.gff")
    seqs, order = mo.read_fasta(fa)
    a = genome.read_gff(gff)
    a.genome = _FakeGenome(_FakeGS(order))
    fl = Flattener(a)
    for k in a.gene:
        fl.add_top(a.gene[k])
    got = []
    for r, name in enumerate(fl.names):
        parts = []
        for s in range(fl.rec_seg_off[r], fl.rec_seg_off[r + 1]):
            piece = seqs[order[fl.seg_contig[s]]][fl.seg_start[s] - 1:fl.seg_end[s]]
            parts.append(mo.reverse_compliment(piece) if fl.seg_strand[s] else piece)
        got.append(">" + name + "\n" + "".join(parts))
    b = mo.read_gff(gff)
    b.genome = seqs
    want = [x for x in mo.annotation_set_get_fasta(b, 'gene').split("\n>")]
    want = [w if w.startswith(">") else ">" + w for w in want if w.strip("\n")]
    want = [w.rstrip("\n") for w in want]
    assert got == want
    # blank-line bookkeeping: genes without CDS-bearing children are tops without entries
    assert sum(1 for t in fl.tops if not t.entries) == 7


def test_shard_bounds_balanced_and_contiguous():
    w = np.array([5, 1, 1, 1, 8, 2, 2, 9, 1, 1], dtype=np.int64)
    for n in (1, 2, 3, 4, 8):
        b = engine.shard_bounds(w, n)
        assert b[0] == 0 and b[-1] == len(w) and len(b) == n + 1
        assert all(b[i] <= b[i + 1] for i in range(n))
    b = engine.shard_bounds(np.ones(1000), 4)
    assert b == [0, 250, 500, 750, 1000]


def test_record_table_slice_roundtrip():
    layout = synth.contig_layout("insect", 200_000, 5)
    ann = synth.synth_annotation(layout, 300, 7)
    tbl = ann.table("cds")
    assert tbl.n_rec == 300 and tbl.rec_seg_off[-1] == tbl.n_seg
    parts = [tbl.slice(b0, b1) for b0, b1 in zip([0, 100, 250], [100, 250, 300])]
    assert sum(p.n_seg for p in parts) == tbl.n_seg
    assert np.array_equal(np.concatenate([p.seg_start for p in parts]), tbl.seg_start)
    assert all(p.rec_seg_off[0] == 0 and p.rec_seg_off[-1] == p.n_seg for p in parts)


def test_synth_annotation_is_consistent():
    layout = synth.contig_layout("human", 3_000_000, 4)
    lens = np.array([l for _, l in layout])
    ann = synth.synth_annotation(layout, 500, 4)
    cnt = np.diff(ann.exon_off)
    ctg = np.repeat(ann.contig, cnt)
    assert (ann.exon_start >= 1).all() and (ann.exon_end >= ann.exon_start).all()
    # emission order: ascending on '+', descending on '-'
    for t in range(0, 500, 37):
        s = ann.exon_start[ann.exon_off[t]:ann.exon_off[t + 1]]
        assert (np.diff(s) > 0).all() if ann.strand[t] == 0 else (np.diff(s) < 0).all()
    cds_cnt = np.diff(ann.cds_off)
    assert (cds_cnt >= 1).all()
    assert ann.spliced_bp("cds") <= ann.spliced_bp("exon")
    assert (ann.exon_end <= lens[ctg] + 10 ** 7).all()


def test_multi_rank_sharding_gloo(tmp_path):
    """world_size-2 gloo: each rank takes its contiguous byte-balanced batch; rank 0 gathers the texts in
    record order (the data path itself has no collective).  The per-rank 'device pass' is replaced by the
    oracle here -- the partition/gather logic is what is under test."""
    import subprocess
    import sys
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    out = tmp_path / "gather.txt"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", script, str(out)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_text() == "OK"


@pytest.mark.parametrize("kind", ["blast", "exonerate"])
def test_aligner_readers_build_the_reference_object_model(kind):
    """read_blast_csv (genome.py:425-499) / read_exonerate + vulgar2gff (genome.py:32-120): IDs in py2 order, coords (incl. the
    lexicographic min/max of vulgar2gff), strands, parents and child lists equal what the reference itself built
    (tests/golden/aligner_model.json, generated by tests/golden/make_golden.py from the shimmed reference)."""
    import json
    gold = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gold, "aligner_model.json")) as fh:
        want = json.load(fh)[kind]
    if kind == "blast":
        aset = genome.read_blast_csv(os.path.join(gold, "aligner_inputs", "obiroi_blast.csv"))
    else:
        aset = genome.read_exonerate(os.path.join(gold, "aligner_inputs", "obiroi_exonerate.txt"))
    got = {}
    for name in aset._dict_names():
        tbl = aset.__dict__[name]
        if tbl:
            got[name] = [[k, v.seqid, list(v.coords) if hasattr(v, "coords") else None, v.strand, v.parent,
                          list(getattr(v, "child_list", []))] for k, v in tbl.items()]
    assert sorted(got) == sorted(want)
    for name in want:
        assert got[name] == want[name], name


def test_lpt_contig_shards():
    """Config-5 sharding rule: contigs longest-first to the least loaded GPU; every contig exactly once, ascending
    inside a share, and -- for the human-scale layout -- the largest share within 3 % of the ideal 1/n."""
    from magot_b200 import orfs
    layout = synth.contig_layout("human", 3_100_000_000, 4)
    lens = [l for _, l in layout]
    for n in (1, 2, 4, 8):
        shards = orfs.lpt_shards(lens, n)
        assert len(shards) == n
        assert sorted(c for sh in shards for c in sh) == list(range(len(lens)))
        assert all(sh == sorted(sh) for sh in shards)
        worst = max(sum(lens[c] for c in sh) for sh in shards)
        assert worst <= 1.03 * sum(lens) / n, (n, worst)
    assert orfs.lpt_shards([], 3) == [[], [], []]


# ---- GFF writers / convert_gff (host-only; genome.py:228-238, :616-645, :733-778; genome_tools.py:527-545)

def _stdout_of(fn, *a, **k):
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a, **k)
    return buf.getvalue()


@pytest.mark.parametrize("args,key", [
    (("minimalGFF3.gff", "gff3", "gtf"), "suite:convert_gff_minimalGFF3_gff3_gtf"),
    (("StandardGTF.gtf", "gtf", "gff3"), "suite:convert_gff_StandardGTF_gtf_gff3"),
    (("StandardGTF.gtf", "gtf", "exon_added_gff3"), "suite:convert_gff_StandardGTF_gtf_exon_added_gff3"),
])
def test_convert_gff_reference_suite_goldens(ref_data, manifest, args, key):
    """The three cksums the reference's own test-suite holds (test_data/test_suite.py:15-17), through the CLI grammar."""
    from cksum import cksum
    from magot_b200 import genome_tools
    out = _stdout_of(genome_tools.main, ["genome_tools.py", "convert_gff", os.path.join(ref_data, args[0]), args[1], args[2]])
    c, n = cksum(out)
    assert {"cksum": c, "bytes": n} == manifest[key]


def test_write_gff_all_formats_match_reference(ref_data, manifest):
    """Every format of get_gff x every annotation fixture against the reference's own output (tests/golden/make_golden.py:
    gff_writer_cases), including the cases where the reference raises (childless parents, gtf without a grand-parent)."""
    from cksum import cksum
    from magot_b200 import genome
    keys = [k for k in manifest if k.startswith("gffwrite:") and "augustus_preset" not in k and "obiroi_exon_based" not in k]
    assert len(keys) == 24
    for k in keys:
        _, f, fmt = k.split(":")
        exp = manifest[k]
        if "raises" in exp:
            with pytest.raises(Exception) as ei:
                genome.write_gff(genome.read_gff(os.path.join(ref_data, f)), fmt)
            assert type(ei.value).__name__ == exp["raises"], k
        else:
            c, n = cksum(genome.write_gff(genome.read_gff(os.path.join(ref_data, f)), fmt))
            assert {"cksum": c, "bytes": n} == exp, k
    ob = os.path.join(ref_data, "O.biroi_NCBIrefseq_gff3Subset.gff")
    for fmt in ("simple gff3", "extended gff3"):
        aset = genome.read_gff(ob, base_features=['exon', 'match_part', 'similarity', 'region'], features_to_ignore=['CDS'])
        c, n = cksum(genome.write_gff(aset, fmt))
        assert {"cksum": c, "bytes": n} == manifest["gffwrite:obiroi_exon_based:%s" % fmt]


def test_read_gff_presets(ref_data, manifest):
    """presets= (genome.py:261-268): 'augustus' equals the reference run with the preset's assignments passed explicitly,
    an unknown name (convert_gff's 'gtf') is a no-op, 'CEGMA' raises like the reference's exec of its malformed list."""
    from cksum import cksum
    from magot_b200 import genome
    gtf = os.path.join(ref_data, "StandardGTF.gtf")
    for fmt in ("simple gff3", "gtf"):
        c, n = cksum(genome.write_gff(genome.read_gff(gtf, presets="augustus"), fmt))
        assert {"cksum": c, "bytes": n} == manifest["gffwrite:StandardGTF.gtf:augustus_preset:%s" % fmt]
    assert genome.write_gff(genome.read_gff(gtf, presets="gtf")) == genome.write_gff(genome.read_gff(gtf))
    with pytest.raises(TypeError):
        genome.read_gff(gtf, presets="CEGMA")


def test_py2_instance_dict_order_matches_oracle_emulator():
    import random
    from py2dict import py2_instance_dict_order
    from magot_b200.py2dict import py2_instance_attr_order
    rnd = random.Random(5)
    pool = ["ID", "seqid", "coords", "feature_type", "annotation_set", "source", "score", "phase", "gene_id", "transcript_id",
            "parent", "strand", "child_list", "Name", "product", "Dbxref", "gbkey", "Note", "gene", "protein_id"]
    for _ in range(200):
        names = rnd.sample(pool, rnd.randint(1, len(pool)))
        for dc in (False, True):
            assert py2_instance_attr_order(names, dc) == py2_instance_dict_order(names, dc)


def test_annotation_set_seqid_helpers_and_cegma(ref_data):
    """AnnotationSet.get_all_seqids / get_seqid (genome.py:550-567), Genome.get_seqids(from_annotations=True) (:935-948) and
    read_cegma_gff (:418-422: the reference's CEGMA preset text cannot be exec'd, TypeError)."""
    a = genome.read_gff(os.path.join(ref_data, "O.biroi_NCBIrefseq_gff3Subset.gff"))
    ids = a.get_all_seqids()
    first_seen = list(dict.fromkeys(o.seqid for n in a._dict_names() for o in a.__dict__[n].values()))
    assert ids == oracle_py2.py2_order(first_seen) and len(ids) == 6
    sub = a.get_seqid(ids[0])
    for n in a._dict_names():
        want = [k for k, o in a.__dict__[n].items() if o.seqid == ids[0]]
        assert sorted(sub.__dict__[n]) == sorted(want)
        assert list(sub.__dict__[n]) == oracle_py2.py2_order(want)
        assert all(sub.__dict__[n][k] is a.__dict__[n][k] for k in want)
    my = genome.Genome.__new__(genome.Genome)
    my.genome_sequence, my.annotations = None, a
    assert my.get_seqids(from_annotations=True) == ids
    with pytest.raises(TypeError):
        genome.read_cegma_gff(os.path.join(ref_data, "StandardGTF.gtf"))


# ---- position_dic (genome.py:981-1100), host half: fill / count from annotations

def _position_dic_fixture():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "position_dic.json")) as fh:
        fx = json.load(fh)
    seqs = {}
    for rec in fx["fasta"].split(">")[1:]:
        head, _, body = rec.partition("\n")
        seqs[head] = body.replace("\n", "")
    return fx, {k: seqs[k] for k in py2dict.py2_order(list(seqs))}         # GenomeSequence iterates in Python-2.7 order


def test_position_dic_fill_and_count_match_reference():
    """fill_from_annotations ("coords" / "start", constant and coords-dependent fill_with) and count_from_annotations against
    the reference's own arrays (tests/golden/make_golden.py:position_dic_cases); no device needed for these."""
    fx, seqs = _position_dic_fixture()
    aset = genome.read_gff(fx["gff"])
    assert list(genome.position_dic(seqs)) == fx["order"]
    bits = lambda pd: {k: "".join("1" if x else "0" for x in v) for k, v in pd.items()}      # noqa: E731
    for feature, fill_type in (("CDS", "coords"), ("gene", "coords"), ("CDS", "start")):
        pd = genome.position_dic(seqs)
        pd.fill_from_annotations(aset, feature, fill_type=fill_type)
        assert bits(pd) == fx["fill"]["%s:%s" % (feature, fill_type)]
    pd = genome.position_dic(seqs, dtype=int)
    pd.fill_from_annotations(aset, "CDS", fill_with="coords[0] % 3")
    assert {k: [int(x) for x in v] for k, v in pd.items()} == fx["fill"]["CDS:coords:int:coords[0] % 3"]
    pd = genome.position_dic(seqs)
    pd.fill_from_annotations(aset, "CDS")
    assert pd.count_from_annotations(aset, "gene") == fx["count"]["gene_over_CDS_fill"]
    # numpy index semantics of the reference's per-position loop: position -1 is the last element, past the end raises
    pd = genome.position_dic({"c": "ACGTACGTAC"})
    b = genome.BaseAnnotation("x", "c", (0, 3), "CDS", None, "+", {}, genome.AnnotationSet())
    fake = genome.AnnotationSet()
    fake.CDS = {"x": b}
    pd.fill_from_annotations(fake, "CDS")
    assert pd["c"].astype(int).tolist() == [1, 1, 1, 0, 0, 0, 0, 0, 0, 1]
    b.coords = (8, 12)
    with pytest.raises(IndexError):
        pd.fill_from_annotations(fake, "CDS")
    assert pd["c"].astype(int).tolist() == [1, 1, 1, 0, 0, 0, 0, 1, 1, 1]


# ---- native GFF reader (csrc/mg_gff.cu) against the oracle's line-by-line restatement on adversarial text -----------------

def _full_signature(tables):
    sig = {}
    for name, tbl in tables.items():
        for k, o in tbl.items():
            d = {a: v for a, v in o.__dict__.items() if a != "annotation_set"}
            sig[(name, k)] = (type(o).__name__[:4], d)
    return sig


def _fuzz_gff(rng, flavour):
    seqids = ["s1", "s2", "chr 3"]
    parent_types = ["gene", "mRNA", "transcript", "match", "five_prime_UTR"]
    base_types = ["exon", "CDS", "region", "match_part"]
    ids = ["g%d" % i for i in range(6)] + ["t%d" % i for i in range(8)] + ["c%d" % i for i in range(5)]
    lines = ["##gff-version 3", "# a comment\twith\ttabs\t1\t2\t3\t4\t5\t6"]
    known = []                                           # IDs of features that can take children
    n_lines = int(rng.integers(5, 160))
    poison_at = int(rng.integers(0, n_lines)) if rng.random() < 0.3 else -1
    for li in range(n_lines):
        r = rng.random()
        if r < 0.04:
            lines.append(rng.choice(["", "#x", "a\tb\tc", "s1\tsrc\tgene\t1\t2\t.\t+\t.\tID=z\textra\tcol", "\r"]))
            continue
        is_parent = rng.random() < 0.45
        ftype = str(rng.choice(parent_types if is_parent else base_types))
        a, b = int(rng.integers(-3, 5000)), int(rng.integers(-3, 5000))
        score = str(rng.choice([".", "1.5", "abc", "1e3", "7", " 2 ", "-0.0", "inf"]))
        strand = str(rng.choice(["+", "-", ".", "?", "+-"]))
        phase = str(rng.choice([".", "0", "1", "2", "3", "01"]))
        attrs = []
        poison = li == poison_at
        if flavour == "gff3":
            my_id = None
            if rng.random() < 0.85:
                my_id = str(rng.choice(ids))
                attrs.append("ID=" + my_id)
            if known and rng.random() < 0.6:
                attrs.append("Parent=" + (str(rng.choice(known)) if not (poison and rng.random() < 0.5) else "nowhere"))
            elif my_id is None:                              # (a line with neither ID nor Parent is filed under None, whose
                my_id = str(rng.choice(ids))                 #  place in a Python-2 dict depends on an address: not generated)
                attrs.append("ID=" + my_id)
            for _k in range(int(rng.integers(0, 4))):
                attrs.append(str(rng.choice(["Name=n1", "Name=n2", "product=a=b=c", "note=x y", "source=over", "strand=*", "Dbxref=", "k=v", ""])))
            if poison and rng.random() < 0.5:
                attrs.append("novalue")
            if my_id is not None and is_parent and ftype != "five_prime_UTR":
                known.append(my_id)
        else:
            t = str(rng.choice(ids[6:14]))
            g = str(rng.choice(ids[:6]))
            style = rng.random()
            if style < 0.8:
                attrs.append('transcript_id "%s"' % t)
                attrs.append(' gene_id "%s"' % g)
            elif style < 0.9:
                attrs.append('gene_id "%s"' % g)
            else:
                attrs.append("gene_id %s" % g)
            for _k in range(int(rng.integers(0, 3))):
                attrs.append(str(rng.choice([' exon_number "2"', ' note "a;b"', ' gene_name "x" "y"', " tag v w", ""])))
            if poison:
                attrs.append(" flag")
        if rng.random() < 0.5:
            rng.shuffle(attrs)
        line = "\t".join([str(rng.choice(seqids)), "src", ftype, str(a), str(b), score, strand, phase, ";".join(attrs)])
        if rng.random() < 0.05:
            line = line.replace("\t", "\r\t", 1) if rng.random() < 0.5 else line + "\r"
        lines.append(line)
    return "\n".join(lines) + ("\n" if rng.random() < 0.8 else "")


@pytest.mark.parametrize("seed", list(range(40)))
def test_native_reader_matches_oracle_on_adversarial_text(seed):
    """Random GFF3 / GTF text with duplicate IDs, repeated and empty attributes, '=' and quotes inside values, CR at odd
    places, comment / blank / short lines, reversed and negative coordinates, odd scores, phases and strands, missing parents:
    the native reader's object model (every instance attribute, table contents and Python-2.7 order) == the oracle's, and
    both stop (None) on the same inputs."""
    import contextlib
    import io
    rng = np.random.default_rng(900 + seed)
    flavour = "gff3" if seed % 2 == 0 else "gtf"
    text = _fuzz_gff(rng, flavour)
    kws = [{}, {"features_to_ignore": ["CDS"], "base_features": ["exon", "region"]},
           {"features_to_ignore": "CDSexon", "features_to_replace": [("mRNA", "transcript"), ("five_prime_UTR", "UTR")]}]
    if flavour == "gtf":
        kws.append({"parents_hierarchy": ["gene_id"], "IDfield": None, "parent_field": None})
    else:
        kws.append({"IDfield": "Name", "parent_field": "Parent"})
    for kw in kws:
        out_a, out_b = io.StringIO(), io.StringIO()
        with contextlib.redirect_stdout(out_a):
            try:
                a = genome.read_gff(text, **kw)
                err_a = None
            except Exception as e:       # noqa: BLE001
                a, err_a = None, type(e).__name__
        with contextlib.redirect_stdout(out_b):
            try:
                b = mo.read_gff(text, **kw)
                err_b = None
            except Exception as e:       # noqa: BLE001
                b, err_b = None, type(e).__name__
        assert err_a == err_b, (kw, err_a, err_b)
        assert (a is None) == (b is None), kw
        if a is None:
            continue
        ta = {name: a.__dict__[name] for name in a._dict_names()}
        tb = {name: b.tables[name] for name in sorted(b.tables)}
        assert {k: list(v) for k, v in ta.items()} == {k: list(v) for k, v in tb.items()}, kw
        assert _full_signature(ta) == _full_signature(tb), kw
