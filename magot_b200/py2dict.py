"""Record order of the reference: CPython 2.7 dict iteration order.

The reference prints records by iterating plain Python-2.7 dicts (AnnotationSet.get_fasta
genome.py:580, exclude_from_fasta genome_tools.py:385, get_genome_fasta genome.py:912) and
read_gff returns a copy.deepcopy of what it built (genome.py:415), which re-inserts every key
in the old table's slot order.  To produce byte-identical files this module replays the
open-addressing table of CPython 2.7 (Objects/dictobject.c: 8 initial slots, probe
i = 5*i + perturb + 1 with perturb >>= 5, grow when fill*3 >= 2*size to the first power of
two > 4*used, or 2*used above 50000 entries) with the 2.7 string hash (hash randomisation
off).  Hashes are computed column-wise with numpy so that 10^5..10^6 IDs take a fraction of
a second; the probe sequence itself is replayed in a tight Python loop over integers.
"""
import numpy as np

_MASK = (1 << 64) - 1


def string_hashes(keys):
    """CPython-2.7 string_hash (64-bit) of every key, as a list of unsigned Python ints."""
    n = len(keys)
    if n == 0:
        return []
    enc = [k.encode("latin-1") if isinstance(k, str) else bytes(k) for k in keys]
    lens = np.fromiter((len(b) for b in enc), dtype=np.int64, count=n)
    maxlen = int(lens.max())
    if maxlen == 0:
        return [0] * n
    buf = np.zeros((n, maxlen), dtype=np.uint8)
    flat = np.frombuffer(b"".join(enc), dtype=np.uint8)
    starts = np.concatenate(([0], np.cumsum(lens)[:-1]))
    rows = np.repeat(np.arange(n), lens)
    cols = np.arange(flat.size) - np.repeat(starts, lens)
    buf[rows, cols] = flat
    with np.errstate(over="ignore"):
        x = buf[:, 0].astype(np.uint64) << np.uint64(7)
        mul = np.uint64(1000003)
        for j in range(maxlen):
            live = lens > j
            nx = (x * mul) ^ buf[:, j].astype(np.uint64)
            x = np.where(live, nx, x)
        x ^= lens.astype(np.uint64)
    x = np.where(lens == 0, np.uint64(0), x)
    x = np.where(x == np.uint64(_MASK), np.uint64(_MASK - 1), x)     # -1 -> -2
    return x.tolist()


def _native_perm(keys, rounds):
    """mg_py2_order (csrc/mg_gff.cu) on latin-1 str keys: permutation as a list, or None when the keys are not all str / the
    library is not loadable (the pure-Python replay below is the same algorithm)."""
    if len(keys) < 64 or not all(type(k) is str for k in keys):
        return None
    try:
        import ctypes
        from ._lib import lib
        enc = "".join(keys).encode("latin-1")
        if len(enc) != sum(map(len, keys)):
            return None
        off = np.zeros(len(keys) + 1, dtype=np.int64)
        np.cumsum(np.fromiter(map(len, keys), dtype=np.int64, count=len(keys)), out=off[1:])
        perm = np.empty(len(keys), dtype=np.int64)
        if lib.mg_py2_order(enc, off.ctypes.data_as(ctypes.c_void_p), len(keys), rounds, perm.ctypes.data_as(ctypes.c_void_p)) != 0:
            return None
        return perm.tolist()
    except Exception:
        return None


def py2_order(keys):
    """Keys (unique, in insertion order) -> the order a CPython 2.7 dict iterates them in."""
    keys = list(keys)
    perm = _native_perm(keys, 1)
    if perm is not None:
        return [keys[i] for i in perm]
    return _py2_order_python(keys)


def _py2_order_python(keys):
    hashes = string_hashes(keys)
    size = 8
    mask = 7
    slots = [-1] * size            # index into keys, -1 = empty
    used = 0
    for idx, h in enumerate(hashes):
        i = h & mask
        if slots[i] != -1:
            perturb = h
            j = i
            while True:
                j = ((j << 2) + j + perturb + 1) & _MASK
                perturb >>= 5
                i = j & mask
                if slots[i] == -1:
                    break
        slots[i] = idx
        used += 1
        if used * 3 >= size * 2:
            minused = (2 if used > 50000 else 4) * used
            newsize = 8
            while newsize <= minused:
                newsize <<= 1
            new = [-1] * newsize
            nmask = newsize - 1
            for k in slots:
                if k != -1:
                    hh = hashes[k]
                    i = hh & nmask
                    if new[i] != -1:
                        perturb = hh
                        j = i
                        while True:
                            j = ((j << 2) + j + perturb + 1) & _MASK
                            perturb >>= 5
                            i = j & nmask
                            if new[i] == -1:
                                break
                    new[i] = k
            slots, size, mask = new, newsize, nmask
    return [keys[k] for k in slots if k != -1]


def py2_order_after_deepcopy(keys):
    """Order after read_gff's copy.deepcopy (genome.py:415): re-insert in old slot order."""
    keys = list(keys)
    perm = _native_perm(keys, 2)
    if perm is not None:
        return [keys[i] for i in perm]
    return _py2_order_python(_py2_order_python(keys))


def py2_update_order(keys):
    """Order of an EMPTY CPython-2.7 dict after `d.update(other)` where `keys` is other's iteration order
    (PyDict_Merge, Objects/dictobject.c): one pre-resize to the first power of two > 2*len(other) when
    len(other)*3 >= 16, then plain insertions with no further growth."""
    keys = list(keys)
    n = len(keys)
    size = 8
    if n * 3 >= 16:
        while size <= 2 * n:
            size <<= 1
    mask = size - 1
    slots = [-1] * size
    for idx, h in enumerate(string_hashes(keys)):
        i = h & mask
        if slots[i] != -1:
            perturb = h
            j = i
            while True:
                j = ((j << 2) + j + perturb + 1) & _MASK
                perturb >>= 5
                i = j & mask
                if slots[i] == -1:
                    break
        slots[i] = idx
    return [keys[k] for k in slots if k != -1]


def py2_instance_attr_order(names, deepcopied=False):
    """Iteration order of an instance __dict__ whose attributes were first assigned in the order `names`.
    copy.deepcopy of an instance (copy.py `_deepcopy_inst` / `_reconstruct`) deep-copies the state dict
    (re-insertion in slot order) and then `y.__dict__.update(state)`."""
    order = py2_order(names)
    if deepcopied:
        order = py2_update_order(py2_order(order))
    return order


def reorder_dict(d, deepcopy=False):
    """Return a new dict with d's items in CPython-2.7 order (keys must be str)."""
    order = py2_order_after_deepcopy(list(d)) if deepcopy else py2_order(list(d))
    return {k: d[k] for k in order}
