"""TEST INFRASTRUCTURE -- runs the reference's OWN code (shimmed by make_ref.py) with the
Python-2.7 dict order restored, so its output equals what `python2 genome_tools.py ...` prints.

Used by tests/golden/make_golden.py (fixture generation), by tests that run in this
container (where /root/reference exists) and by `bench.py --impl reference`.  On a box
without /root/reference it works only if the git-ignored oracle/_ref/ travelled there.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref                      # noqa: E402
from py2dict import py2_order, py2_order_after_deepcopy, py2_instance_dict_order   # noqa: E402

_genome = None


def ref():
    """The shimmed reference `genome` module, or None."""
    global _genome
    if _genome is None:
        _genome = make_ref.load()
    return _genome


def reorder_annotation_set(aset):
    """Apply CPython-2.7 + deepcopy (genome.py:415) iteration order to every feature dict."""
    for name, val in list(aset.__dict__.items()):
        if type(val) == dict:
            order = py2_order_after_deepcopy(list(val))
            aset.__dict__[name] = {k: val[k] for k in order}
    return aset


def reorder_genome_sequence(gs):
    """Apply CPython-2.7 iteration order to a GenomeSequence (no deepcopy on that path)."""
    items = {k: gs[k] for k in py2_order(list(gs))}
    gs.clear()
    gs.update(items)
    return gs


def load_genome(fasta, gff=None, truncate_names=False, **read_gff_kwargs):
    g = ref()
    my = g.Genome(fasta, truncate_names=truncate_names)
    if my.genome_sequence is not None:
        reorder_genome_sequence(my.genome_sequence)
    if gff is not None:
        my.read_gff(gff, **read_gff_kwargs)
        if my.annotations is not None:
            reorder_annotation_set(my.annotations)
    return my


def gff2fasta(fasta, gff, from_exons="False", seq_type="nucleotide", longest="False", genomic="False"):
    """stdout of genome_tools.py:324-330."""
    if from_exons == "True":
        my = load_genome(fasta, gff, features_to_ignore="CDS", features_to_replace=[('exon', 'CDS')])
    else:
        my = load_genome(fasta, gff)
    return my.annotations.get_fasta('gene', seq_type=seq_type, longest=eval(longest), genomic=eval(genomic)) + "\n"


def reorder_instance_dicts(aset, deepcopied=True):
    """Give the set's own __dict__ and every annotation object's __dict__ the iteration order Python 2.7 leaves them
    in (write_gff genome.py:230 iterates the former, "extended gff3" genome.py:631 / :751 the latter)."""
    objs = [aset]
    for val in aset.__dict__.values():
        if type(val) == dict:
            objs.extend(val.values())
    for o in objs:
        d = o.__dict__
        order = py2_instance_dict_order(list(d), deepcopied)
        new = {k: d[k] for k in order}
        d.clear()
        d.update(new)
    return aset


def write_gff(gff, gff_format="simple gff3", **read_gff_kwargs):
    """genome.write_gff(genome.read_gff(gff, ...), gff_format) of the reference, Python-2.7 orders restored."""
    g = ref()
    aset = g.read_gff(gff, **read_gff_kwargs)
    reorder_annotation_set(aset)
    reorder_instance_dicts(aset, deepcopied=True)
    return g.write_gff(aset, gff_format)


def convert_gff(gff, input_format, output_format):
    """stdout of genome_tools.py:527-545 (input formats that are not exec-presets only: the shim cannot rebind locals)."""
    fmt = {"gff3": "simple gff3", "gtf": "gtf", "exon_added_gff3": "exon added gff3"}[output_format]
    assert input_format not in ("augustus", "RepeatMasker", "CEGMA")
    return write_gff(gff, fmt) + "\n"


def reorder_no_deepcopy(aset):
    """CPython-2.7 iteration order of dicts that were filled in place (read_blast_csv / read_exonerate, no deepcopy)."""
    for name, val in list(aset.__dict__.items()):
        if type(val) == dict:
            aset.__dict__[name] = {k: val[k] for k in py2_order(list(val))}
    return aset


def _aligner2fasta(fasta, reader, path):
    g = ref()
    my = g.Genome(fasta)
    reorder_genome_sequence(my.genome_sequence)
    getattr(my, reader)(path)
    reorder_no_deepcopy(my.annotations)
    out = [my.annotations.match[m].get_fasta() for m in my.annotations.match]
    return my, "\n".join(out) + "\n"


def blast_csv2fasta(fasta, blast_csv):
    """stdout of genome_tools.py:265-271."""
    return _aligner2fasta(fasta, "read_blast_csv", blast_csv)[1]


def exonerate2fasta(fasta, exonerate_file):
    """stdout of genome_tools.py:274-280."""
    return _aligner2fasta(fasta, "read_exonerate", exonerate_file)[1]


def aligner_model(fasta, reader, path):
    """Object model the reference builds (IDs in py2 order, coords, strand, parent, children) as plain data."""
    my = _aligner2fasta(fasta, reader, path)[0]
    model = {}
    for name, val in my.annotations.__dict__.items():
        if type(val) == dict and val:
            model[name] = [[k, v.seqid, list(v.get_coords()) if hasattr(v, "coords") else None, v.strand, v.parent,
                            list(getattr(v, "child_list", []))] for k, v in val.items()]
    return model


def mask_from_gff(fasta, gff, **kw):
    """stdout of genome_tools.py:394-428 (the reference's own function), records put back into Python-2.7 dict order."""
    import contextlib
    import importlib
    import io
    ref()
    gt = importlib.import_module("genome_tools")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gt.mask_from_gff(fasta, gff, **kw)
    text = buf.getvalue()
    if not text.startswith(">"):
        return text
    recs = {}
    lines = text.split("\n")
    for k in range(0, len(lines) - 1, 2):
        recs[lines[k][1:]] = lines[k + 1]
    return "".join(">" + n + "\n" + recs[n] + "\n" for n in py2_order(list(recs)))
