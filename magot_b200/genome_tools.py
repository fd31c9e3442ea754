#!/usr/bin/env python3
"""Drop-in mirror of the sequence entry points of MAGOT's `genome_tools` script
(reference: /root/reference/genome_tools.py).

Usage is unchanged:  python -m magot_b200.genome_tools <function> [positional ...] [key=value ...]
Every argument arrives as a string ("True"/"False" are interpreted inside the tools, as in the
reference) and every tool prints to stdout exactly what the reference prints.

Kept (file:line in genome_tools.py): gff2fasta :324, cds2pep :664, coords2fasta :656,
get_seq_from_fasta :483, exclude_from_fasta :377, extract_upstream_downstream :457,
blast_csv2fasta :265, exonerate2fasta :274, mask_from_gff :394, convert_gff :527,
dna2orfs :145 (the reference's version cannot run -- it calls str.translate with keyword
arguments -- this one does what it intended, on the device).  The dispatcher (main :25-45)
keeps the grammar but looks the function up in a table instead of eval()-ing a string.
"""
import sys

import numpy as np

from . import genome
from . import engine


def _truth(s):
    """eval("True") / eval("False") of the reference, without eval."""
    if s in ("True", "False"):
        return s == "True"
    if isinstance(s, bool):
        return s
    raise NameError("name %r is not defined" % (s,))


def gff2fasta(genome_sequence, gff, from_exons="False", seq_type="nucleotide", longest="False", genomic="False"):
    """genome_tools.py:324-330."""
    my_genome = genome.Genome(genome_sequence)
    if from_exons == "True":
        # the reference passes features_to_ignore as the *string* "CDS" (substring test) after renaming
        # exon -> CDS, so every exon and CDS line is ignored; reproduced as is.
        my_genome.read_gff(gff, features_to_ignore="CDS", features_to_replace=[('exon', 'CDS')])
    else:
        my_genome.read_gff(gff)
    print(my_genome.annotations.get_fasta('gene', seq_type=seq_type, longest=_truth(longest), genomic=_truth(genomic)))


def convert_gff(gff, input_format, output_format):
    """genome_tools.py:527-545 -- re-write an annotation file: input `gff3` or a read_gff preset name (anything else, e.g.
    `gtf`, reads with the defaults), output `gff3`, `gtf` or `exon_added_gff3`.  Host-only (no sequence is touched)."""
    presets = None if input_format == 'gff3' else input_format
    if output_format == 'gff3':
        gff_format = "simple gff3"
    elif output_format == 'gtf':
        gff_format = 'gtf'
    elif output_format == "exon_added_gff3":
        gff_format = "exon added gff3"
    else:
        print("currently only writes 'gff3' and 'gtf' format")
        return None
    annotations = genome.read_gff(gff, presets=presets)
    print(genome.write_gff(annotations, gff_format))


def blast_csv2fasta(genome_sequence, blast_csv):
    """genome_tools.py:265-271 -- one record per blast hit (`match`), all hits through ONE device plan."""
    my_genome = genome.Genome(genome_sequence)
    my_genome.read_blast_csv(blast_csv)
    print(my_genome.annotations.get_fasta('match'))


def exonerate2fasta(genome_sequence, exonerate_file):
    """genome_tools.py:274-280 -- one record per exonerate alignment (`match`), its match_parts spliced."""
    my_genome = genome.Genome(genome_sequence)
    my_genome.read_exonerate(exonerate_file)
    print(my_genome.annotations.get_fasta('match'))


def mask_from_gff(genome_sequence, gff, mask_type="soft", overwrite_softmask="True", feature_type="CDS"):
    """genome_tools.py:394-428 -- soft- (lower-case) or hard- ('N') mask every `feature_type` interval of a GFF and print
    the genome, one line per sequence.  The intervals are scattered onto the packed genome on the device (K5).
    Kept from the reference: seqid = first word of the header, a repeated header starts the sequence over, any GFF line with
    more than 5 tabs counts, coordinates follow Python slice rules, the printed order is the reference's dict order.
    Not kept: a hard mask whose slice is clamped changes the sequence LENGTH in the reference (list slice assignment);
    that raises NotImplementedError here."""
    with open(genome_sequence, "rb") as fh:
        data = fh.read()
    if overwrite_softmask in ("True", "T", "true", "t", "TRUE"):
        upper_first = True
    elif overwrite_softmask in ("False", "F", "false", "f", "FALSE"):
        upper_first = False
    else:
        upper_first = None
    seqs = {}
    working = None
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                                     # no line after the final newline
    for line in lines:
        if line[:1] == b">":
            working = line.decode("latin-1").split()[0][1:].replace('\r', '')
            seqs[working] = []
        elif line or working is None:
            if working is None:
                raise KeyError("")                      # sequence before the first header: genome_dict[""] in the reference
            if upper_first is None:
                print("Invalid option for 'overwrite_softmask', argument accepts 'True' or 'False'")
                return None
            seqs[working].append(line.replace(b"\r", b""))
    names = genome._order(list(seqs))
    arrays = [np.frombuffer(b"".join(seqs[n]), dtype=np.uint8) for n in names]
    index = {n: i for i, n in enumerate(names)}
    cid, lo, hi = [], [], []
    with open(gff, encoding="latin-1", newline="\n") as fh:
        for line in fh:
            if line.count('\t') > 5:
                fields = line.split('\t')
                if fields[2] == feature_type:
                    start, stop = int(fields[3]), int(fields[4])
                    ci = index[fields[0]]                # KeyError for an unknown seqid, as the reference
                    a, b, _ = slice(start - 1, stop).indices(arrays[ci].size)
                    b = max(a, b)
                    if mask_type == "soft":
                        pass
                    elif mask_type == "hard":
                        if b - a != max(1 + stop - start, 0):
                            raise NotImplementedError("hard mask %s:%d-%d is clamped by the sequence end: the reference would "
                                                      "change the sequence length here" % (fields[0], start, stop))
                    else:
                        print("Invalid option for mask_type, argument accepts 'soft' and 'hard'")
                        return None
                    cid.append(ci)
                    lo.append(a)
                    hi.append(b)
    g = engine.DeviceGenome([a.size for a in arrays], device=0)
    try:
        for ci, a in enumerate(arrays):
            g.pack(ci, a)
        g.finalize()
        g.mask(cid, lo, hi, hard=(mask_type == "hard"), upper_first=bool(upper_first))
        for ci, n in enumerate(names):
            L = arrays[ci].size
            print(">" + n + "\n" + (g.fetch(ci, 0, L).decode("latin-1") if L else ""))
    finally:
        g.close()


def cds2pep(fasta_file):
    """genome_tools.py:664-675 -- all records are translated in ONE device launch."""
    headers = []
    seqs = []
    working = []
    have = False
    order = []                       # ('h', idx) / ('s', idx) in print order
    with open(fasta_file, encoding="latin-1", newline="\n") as fh:
        for original_line in fh:
            line = original_line.replace('\n', '').replace('\r', '')
            if line[0] == '>':
                if have:
                    seqs.append("".join(working).encode("latin-1"))
                    order.append(('s', len(seqs) - 1))
                    working = []
                    have = False
                headers.append(line)
                order.append(('h', len(headers) - 1))
            else:
                working.append(line)
                if line != "":
                    have = True
    seqs.append("".join(working).encode("latin-1"))
    order.append(('s', len(seqs) - 1))
    peps = genome._translate_many(seqs)
    out = []
    for kind, idx in order:
        out.append(headers[idx] if kind == 'h' else str(peps[idx]))
    sys.stdout.write("\n".join(out) + "\n")


def coords2fasta(fasta_file, seqid, start, stop, truncate_names="False"):
    """genome_tools.py:656-661: 1-based inclusive coordinates, Python slice clamping."""
    print(">" + seqid + ":" + start + "-" + stop)
    print(genome.Genome(fasta_file, truncate_names=_truth(truncate_names)).genome_sequence[seqid][int(start) - 1:int(stop)])


def get_seq_from_fasta(genome_sequence, seq_name, truncate_names="False"):
    """genome_tools.py:483-485."""
    my_genome = genome.Genome(genome_sequence, truncate_names=_truth(truncate_names))
    print(my_genome.get_scaffold_fasta(seq_name))


def exclude_from_fasta(fasta, exclude_list, just_firstword="False"):
    """genome_tools.py:377-391 (records in the reference's dict order)."""
    my_fasta = genome.Genome(fasta)
    try:
        exlist = open(exclude_list).read().replace('\r', '').split('\n')
    except Exception:
        exlist = exclude_list.split(',')
    for seqid in my_fasta.genome_sequence:
        seqid_fixed = seqid.split()[0] if just_firstword == "True" else seqid
        if seqid_fixed not in exlist:
            print('>' + seqid + '\n' + my_fasta.genome_sequence[seqid])


def extract_upstream_downstream(genome_sequence, gff, sequence_length, stream, feature_type="gene", namefrom="ID",
                                truncate_names="True"):
    """genome_tools.py:457-480.  All flanks are fetched in one device plan; the slice arithmetic
    (including Python's negative-index semantics for flanks that run off the contig start) is the
    device's, via start-1/end == Python slice bounds."""
    sequence_dict = genome.GenomeSequence(genome_sequence, truncate_names=_truth(truncate_names))
    n = int(sequence_length)
    names, segs = [], []
    have_sequence = False
    last_seg = None
    with open(gff, encoding="latin-1", newline="\n") as fh:
        for line in fh:
            if line.count('\t') > 5 and line[0] != "#":
                fields = line.split('\t')
                if fields[2] == feature_type:
                    name = None
                    coords = sorted([int(fields[3]), int(fields[4])])
                    for attribute in fields[-1].split(';'):
                        if namefrom == attribute.split('=')[0]:
                            name = attribute.split('=')[1].replace('\r', '').replace('\n', '')
                    ci = sequence_dict.contig_index(fields[0])
                    if (stream == "up" and fields[6] == "+") or (stream == "down" and fields[6] == "-"):
                        stop = coords[0] - 1
                        last_seg = (ci, stop - n + 1, stop, 0)          # contig[stop-n:stop]
                        have_sequence = True
                    elif (stream == "down" and fields[6] == "+") or (stream == "up" and fields[6] == "-"):
                        start = coords[1]
                        last_seg = (ci, start + 1, start + n, 1)        # rc(contig[start:start+n])
                        have_sequence = True
                    if not have_sequence:
                        raise UnboundLocalError("local variable 'sequence' referenced before assignment")
                    # like the reference, a feature with another strand value re-uses the previous `sequence`
                    names.append(name)                                   # None -> 'seq<kept so far>' below
                    segs.append(last_seg)
    if not names:
        print("")
        return
    R = len(names)
    zero = np.zeros(R, dtype=np.int32)
    tbl = engine.RecordTable(np.arange(R + 1), [s[0] for s in segs], [s[1] for s in segs], [s[2] for s in segs],
                             [s[3] for s in segs], np.zeros(R, dtype=np.int64), zero, zero, np.zeros(0, dtype=np.uint8))
    text, (nuc_len, _) = sequence_dict._engine().run_table(tbl, want_lengths=True)
    off = np.concatenate(([0], np.cumsum(nuc_len)))
    output_seqs = []
    for k in range(R):
        if nuc_len[k] == n:                                              # genome_tools.py:478
            name = names[k] if names[k] is not None else 'seq' + str(len(output_seqs))
            output_seqs.append('>' + name + '\n' + text[off[k]:off[k + 1]].decode("latin-1"))
    print("\n".join(output_seqs))


def dna2orfs(fasta_location, output_file, from_atg=False, longest=False, min_orf="0"):
    """What genome_tools.py:145-180 intended (the reference's own version raises TypeError because
    GenomeSequence values are plain str): six-frame translation of every contig, split on stops,
    written as '>{seqid}-pos:{orf_start}' records, or one '>{seqid}_longestORF' record per contig.
    `min_orf` (residues, new) filters short ORFs on the device; "0" reproduces the reference list."""
    from .orfs import contig_orfs
    dna = genome.Genome(fasta_location)
    gs = dna.genome_sequence
    with open(output_file, 'w', encoding="latin-1", newline="\n") as out:
        for seq in gs:
            L = len(gs[seq])
            recs, aa = contig_orfs(gs, seq, int(min_orf))
            candidate = None
            longest_orf_len = 0
            for r, orf in zip(recs, aa):
                strand_plus = not r["minus"]
                frame = int(r["frame"])
                orf_start = (frame if strand_plus else L - frame) + 3 * int(r["start"])
                if from_atg:
                    output_orf = 'M' + ''.join(orf.split('M')[1:])
                else:
                    output_orf = orf
                if longest:
                    if len(output_orf) > longest_orf_len:
                        candidate = '>' + seq + '_longestORF\n' + output_orf + '\n'
                        longest_orf_len = len(output_orf)
                else:
                    out.write('>' + seq + '-pos:' + str(orf_start) + '\n' + output_orf + '\n')
            if longest:
                out.write(candidate)        # TypeError when the contig has no ORF at all (IndexError in the reference)


FUNCTIONS = {f.__name__: f for f in (gff2fasta, cds2pep, coords2fasta, get_seq_from_fasta, exclude_from_fasta,
                                     extract_upstream_downstream, dna2orfs, blast_csv2fasta, exonerate2fasta, mask_from_gff,
                                     convert_gff)}


def help_func():
    print("\ngenome_tools script from MAGOT (B200 path).\n\nUsage: " + sys.argv[0] +
          " function [option1=<option1 choice> ...] \n\nFunctions:\n    " + '\n    '.join(sorted(FUNCTIONS)))


def main(argv=None):
    """genome_tools.py:25-45 -- same grammar: `name=value` -> keyword argument (value = text between the
    first and second '='), anything else positional; all values are strings."""
    argv = sys.argv if argv is None else argv
    program = argv[1]
    arguments = argv[2:]
    if program in ('-h', '-help', '--help', 'help'):
        help_func()
        return None
    if len(arguments) > 0 and arguments[0] in ('-h', '--help', '-help', 'help', '--h'):
        print('hey')
        return None
    if program not in FUNCTIONS:
        raise NameError("name %r is not defined" % program)
    args, kwargs = [], {}
    for argument in arguments:
        if '=' in argument:
            argsplit = argument.split('=')
            kwargs[argsplit[0]] = argsplit[1]
        else:
            args.append(argument)
    return FUNCTIONS[program](*args, **kwargs)


if __name__ == "__main__":
    main()
