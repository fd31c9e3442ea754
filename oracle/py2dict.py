"""TEST INFRASTRUCTURE -- CPython 2.7 `dict` iteration-order emulator (oracle side).

The reference emits records in the iteration order of plain Python-2.7 dicts
(`AnnotationSet.get_fasta`, genome.py:580; `exclude_from_fasta`,
genome_tools.py:385; `Genome.get_genome_fasta`, genome.py:912) and `read_gff`
returns a `copy.deepcopy` of the set it built (genome.py:415), which re-inserts
every key of every dict in the *old* dict's slot order.  Running the shimmed
reference under Python 3 yields insertion order instead, so the oracle applies
this emulator to recover the order the real (Python 2.7) reference prints.

Restated from CPython 2.7 Objects/dictobject.c (lookdict_string, insertdict,
dict_set_item_by_hash_or_entry, dictresize) and Objects/stringobject.c
(string_hash) with hash randomisation off (the 2.7 default).  Deliberately the
slow, literal version: one Python object per slot, no vectorisation.  Pinned by
tests/test_oracle_golden.py against the reference's own goldens
(test_data/test_suite.py:8 and :12).
"""

MASK64 = (1 << 64) - 1
PERTURB_SHIFT = 5
MINSIZE = 8


def py2_string_hash(s):
    """stringobject.c:string_hash for a byte string (64-bit `long`), returned unsigned."""
    if isinstance(s, str):
        s = s.encode("latin-1")
    n = len(s)
    if n == 0:
        return 0
    x = (s[0] << 7) & MASK64
    for c in s:
        x = ((1000003 * x) & MASK64) ^ c
    x ^= n
    if x == MASK64:          # -1 -> -2
        x = MASK64 - 1
    return x


class Py2Dict(object):
    """Insert-only model of a CPython 2.7 dict keyed by byte strings."""

    def __init__(self):
        self.mask = MINSIZE - 1
        self.slots = [None] * MINSIZE      # entries are (hash, key)
        self.used = 0                      # == fill: nothing is ever deleted
        self.keyset = set()

    def _insert_clean(self, slots, mask, h, key):
        i = h & mask
        perturb = h
        while slots[i & mask] is not None:
            i = ((i << 2) + i + perturb + 1) & MASK64
            perturb >>= PERTURB_SHIFT
        slots[i & mask] = (h, key)

    def _resize(self, minused):
        newsize = MINSIZE
        while newsize <= minused:
            newsize <<= 1
        new = [None] * newsize
        for e in self.slots:
            if e is not None:
                self._insert_clean(new, newsize - 1, e[0], e[1])
        self.slots = new
        self.mask = newsize - 1

    def insert(self, key):
        if key in self.keyset:             # value replaced in place: no order change
            return
        self.keyset.add(key)
        self._insert_clean(self.slots, self.mask, py2_string_hash(key), key)
        self.used += 1
        if self.used * 3 >= (self.mask + 1) * 2:
            self._resize((2 if self.used > 50000 else 4) * self.used)

    def keys(self):
        return [e[1] for e in self.slots if e is not None]


def py2_order(keys):
    """Iteration order of a py2 dict into which `keys` were inserted in the given order."""
    d = Py2Dict()
    for k in keys:
        d.insert(k)
    return d.keys()


def py2_order_after_deepcopy(keys):
    """Order after `copy.deepcopy` (genome.py:415): keys re-inserted in old slot order."""
    return py2_order(py2_order(keys))


def py2_update_order(keys):
    """Order of an empty py2 dict after `d.update(other)`, `keys` = other's iteration order.  dictobject.c:PyDict_Merge
    grows the target ONCE up front (to the first power of two > 2*len(other), when len(other)*3 >= 2*8) and then
    calls insertdict per entry, which never resizes."""
    d = Py2Dict()
    n = len(keys)
    if n * 3 >= MINSIZE * 2:
        d._resize(2 * n)
    for k in keys:
        d._insert_clean(d.slots, d.mask, py2_string_hash(k), k)
    return d.keys()


def py2_instance_dict_order(names, deepcopied):
    """Iteration order of an old-style instance's __dict__ (attributes first set in the order `names`).
    copy.py:_deepcopy_inst deep-copies the dict (`_deepcopy_dict`: re-insertion in slot order) and then does
    `y.__dict__.update(state)`."""
    order = py2_order(names)
    if deepcopied:
        order = py2_update_order(py2_order(order))
    return order
