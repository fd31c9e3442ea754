"""ctypes binding of the native GFF3 / GTF reader + flattener (csrc/mg_gff.cu, host code inside libmagot_b200.so).

`parse()` turns annotation text into an integer model (every string interned once; ID naming, de-dup, implicit parents
and child lists done on ids -- read_gff, genome.py:242-415); `Model.flatten()` turns one feature table of that model
into the SoA interval tables K1 consumes (AnnotationSet.get_fasta / ParentAnnotation.get_fasta, genome.py:578-582,
:677-731).  magot_b200.genome builds the reference's Python objects from the model only when somebody asks for them.
"""
import ctypes
import struct

import numpy as np

from ._lib import lib, check
from .tables import RecordTable

_DT = {1: np.int8, 4: np.int32, 8: np.int64}
_FLOAT_COLS = ("score",)

STATUS_OK, GFF2_NO_VALUE, MISSING_PARENT, GFF3_NO_EQ, BAD_INT, GFF2_NO_KEY, PARENT_IS_BASE, ATTR_CLASH, FEW_FIELDS, NONE_TWICE = range(10)


def _s(x):
    b = x.encode("latin-1") if isinstance(x, str) else bytes(x)
    return struct.pack("<i", len(b)) + b


def _names(v):
    """(is_str, [strings]) of a features_to_ignore / base_features argument: a str is tested by substring (`x in "CDS"`)."""
    if isinstance(v, str):
        return 1, [v]
    return 0, [x for x in v if isinstance(x, str)]


def pack_opts(version, features_to_ignore, base_features, parents_hierarchy, features_to_replace, IDfield, parent_field,
              table_names=(), nondict_names=(), existing=()):
    """The option blob mg_gff_parse reads.  existing: [(table index, ID, is_parent, [children])]."""
    out = [struct.pack("<i", version)]
    for v in (features_to_ignore, base_features):
        is_str, names = _names(v)
        out.append(struct.pack("<ii", is_str, len(names)))
        out.extend(_s(x) for x in names)
    hier = [x for x in parents_hierarchy]
    out.append(struct.pack("<i", len(hier)))
    out.extend(_s(x) for x in hier)
    out.append(struct.pack("<i", len(features_to_replace)))
    for a, b in features_to_replace:
        out.append(_s(a) + _s(b))
    out.append(struct.pack("<i", 0 if IDfield is None else 1) + _s(IDfield or ""))
    out.append(struct.pack("<i", 0 if parent_field is None else 1) + _s(parent_field or ""))
    out.append(struct.pack("<i", len(table_names)))
    out.extend(_s(x) for x in table_names)
    out.append(struct.pack("<i", len(nondict_names)))
    out.extend(_s(x) for x in nondict_names)
    out.append(struct.pack("<i", len(existing)))
    for t, ID, is_parent, children in existing:
        out.append(struct.pack("<i", t) + _s(ID) + struct.pack("<ii", 1 if is_parent else 0, len(children)))
        out.extend(_s(c) for c in children)
    return b"".join(out)


class Model(object):
    """Integer model of one read_gff call (mg_gff handle).  Keeps the text alive: the model's strings point into it."""

    def __init__(self, text, opts):
        self.text = text
        self.handle = ctypes.c_void_p()
        buf = (ctypes.c_char * len(text)).from_buffer_copy(text) if not isinstance(text, bytes) else text
        self._buf = buf
        check(lib.mg_gff_parse(buf, len(text), opts, len(opts), ctypes.byref(self.handle)))
        info = (ctypes.c_int64 * 16)()
        check(lib.mg_gff_info(self.handle, info))
        (self.status, self.err_a, self.err_b, self.n_rows, self.n_strings, self.n_attr, self.n_child, self.n_tables,
         self.n_lines, self.has_id_field, self.has_parent_field, self.none_id) = [int(x) for x in info[:12]]
        self._cols = {}
        self._all = None

    def column(self, name):
        c = self._cols.get(name)
        if c is None:
            ptr, n, el = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
            check(lib.mg_gff_column(self.handle, name.encode(), ctypes.byref(ptr), ctypes.byref(n), ctypes.byref(el)))
            dt = np.float64 if name in _FLOAT_COLS else _DT[el.value]
            if n.value == 0:
                c = np.zeros(0, dtype=dt)
            else:
                c = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(n.value * el.value,)).view(dt)
            self._cols[name] = c
        return c

    def strings(self, ids):
        """Python strs (latin-1) of the given string ids."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        off = np.zeros(ids.size + 1, dtype=np.int64)
        check(lib.mg_gff_strings(self.handle, ids.ctypes.data_as(ctypes.c_void_p), ids.size, None, 0, off.ctypes.data_as(ctypes.c_void_p)))
        pool = np.empty(max(int(off[-1]), 1), dtype=np.uint8)
        check(lib.mg_gff_strings(self.handle, ids.ctypes.data_as(ctypes.c_void_p), ids.size, pool.ctypes.data_as(ctypes.c_void_p), pool.size,
                                 off.ctypes.data_as(ctypes.c_void_p)))
        text = pool.tobytes().decode("latin-1")
        o = off.tolist()
        return [text[o[i]:o[i + 1]] for i in range(ids.size)]

    def all_strings(self):
        if self._all is None:
            self._all = self.strings(np.arange(self.n_strings, dtype=np.int32))
            if 0 <= self.none_id < len(self._all):
                self._all[self.none_id] = None
        return self._all

    def string(self, sid):
        return None if sid < 0 or sid == self.none_id else self.strings([sid])[0]

    def py2_order(self, ids, deepcopy=False):
        """Indices into `ids` (string ids, unique) in the order a CPython-2.7 dict keyed by those strings iterates them
        (after a deepcopy: re-inserted once in slot order), computed on the model's own strings (mg_gff_py2_order)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        perm = np.empty(ids.size, dtype=np.int64)
        check(lib.mg_gff_py2_order(self.handle, ids.ctypes.data_as(ctypes.c_void_p), ids.size, 2 if deepcopy else 1,
                                   perm.ctypes.data_as(ctypes.c_void_p)))
        return perm

    def find(self, s):
        b = s.encode("latin-1")
        return int(lib.mg_gff_find(self.handle, b, len(b)))

    def table_rows(self, name):
        """Rows of the table `name` in dict insertion order, or None when the model has no such table."""
        sid = self.find(name)
        names = self.column("table_name")
        hit = np.nonzero(names == sid)[0] if sid >= 0 else []
        if len(hit) == 0:
            return None
        off = self.column("table_off")
        t = int(hit[0])
        return self.column("table_rows")[int(off[t]):int(off[t + 1])]

    def flatten(self, tops, contig_of, framing=True):
        """(RecordTable, top_rec_off, rec_name ids) of the given top rows, or (None, status, err id) when the model needs the
        object path (mixed children, unknown child or seqid, odd strand)."""
        tops = np.ascontiguousarray(tops, dtype=np.int64)
        contig_of = np.ascontiguousarray(contig_of, dtype=np.int32)
        h = ctypes.c_void_p()
        check(lib.mg_gff_flatten(self.handle, tops.ctypes.data_as(ctypes.c_void_p), tops.size, contig_of.ctypes.data_as(ctypes.c_void_p),
                                 contig_of.size, 1 if framing else 0, ctypes.byref(h)))
        try:
            def col(name, dt):
                ptr, n, el = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
                check(lib.mg_gff_flat_column(h, name.encode(), ctypes.byref(ptr), ctypes.byref(n), ctypes.byref(el)))
                if n.value == 0:
                    return np.zeros(0, dtype=dt)
                return np.array(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(n.value * el.value,)).view(dt))
            st = col("status", np.int64)
            if st[0] != 0:
                return None, int(st[0]), int(st[1])
            tbl = RecordTable(col("rec_seg_off", np.int64), col("seg_contig", np.int32), col("seg_start", np.int64), col("seg_end", np.int64),
                              col("seg_strand", np.int8), col("rec_lit_off", np.int64), col("rec_pre", np.int32), col("rec_suf", np.int32),
                              col("lit", np.uint8), col("rec_phase", np.int8))
            return tbl, col("top_rec_off", np.int64), col("rec_name", np.int32)
        finally:
            lib.mg_gff_flat_destroy(h)

    def close(self):
        if self.handle:
            lib.mg_gff_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
