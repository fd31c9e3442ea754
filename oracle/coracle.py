"""TEST INFRASTRUCTURE -- ctypes access to the C restatement (oracle/oracle.c -> oracle/liboracle.so).

Used by tests (checker at sizes where the Python restatement is too slow) and by bench.py's
single-threaded "port" cpu_baseline leg.  Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")


def build():
    src = os.path.join(HERE, "oracle.c")
    if not os.path.isfile(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", SO, src])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.mo_translate.restype = ctypes.c_int64
        _lib.mo_translate_fwd.restype = ctypes.c_int64
        _lib.mo_splice.restype = ctypes.c_int64
        _lib.mo_splice_translate.restype = ctypes.c_int64
        _lib.mo_sixframe.restype = ctypes.c_int64
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def revcomp(seq):
    a = np.frombuffer(seq, dtype=np.uint8)
    out = np.empty(a.size, dtype=np.uint8)
    lib().mo_revcomp(_p(a), ctypes.c_int64(a.size), _p(out))
    return out.tobytes()


def translate(seq, frame=0, minus=False, trimX=True):
    a = np.frombuffer(seq, dtype=np.uint8)
    tmp = np.empty(max(a.size, 1), dtype=np.uint8)
    out = np.empty(a.size // 3 + 4, dtype=np.uint8)
    m = lib().mo_translate(_p(a), ctypes.c_int64(a.size), frame, int(minus), int(trimX), _p(tmp), _p(out))
    return None if m < 0 else out[:m].tobytes()


def splice(contigs, rec_off, cid, lo, hi, minus):
    """contigs: list of bytes; intervals 0-based half-open, already clamped. Returns (text, offsets)."""
    keep = [np.frombuffer(c, dtype=np.uint8) for c in contigs]
    ptrs = (ctypes.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
    rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
    cid = np.ascontiguousarray(cid, dtype=np.int32)
    lo = np.ascontiguousarray(lo, dtype=np.int64)
    hi = np.ascontiguousarray(hi, dtype=np.int64)
    minus = np.ascontiguousarray(minus, dtype=np.int8)
    total = int(np.maximum(hi - lo, 0).sum())
    out = np.empty(max(total, 1), dtype=np.uint8)
    n_rec = rec_off.size - 1
    out_off = np.zeros(n_rec + 1, dtype=np.int64)
    w = lib().mo_splice(ptrs, ctypes.c_int64(n_rec), _p(rec_off), _p(cid), _p(lo), _p(hi), _p(minus), _p(out), _p(out_off))
    return out[:w], out_off


def splice_translate(nuc, nuc_off):
    nuc = np.ascontiguousarray(nuc, dtype=np.uint8)
    nuc_off = np.ascontiguousarray(nuc_off, dtype=np.int64)
    n_rec = nuc_off.size - 1
    aa = np.empty(nuc.size // 3 + n_rec + 4, dtype=np.uint8)
    aa_off = np.zeros(n_rec + 1, dtype=np.int64)
    aa_len = np.zeros(n_rec, dtype=np.int64)
    w = lib().mo_splice_translate(_p(nuc), ctypes.c_int64(n_rec), _p(nuc_off), _p(aa), _p(aa_off), _p(aa_len))
    return aa[:w], aa_off, aa_len


def sixframe(seq, min_aa=0):
    """Returns (aa bytes back to back, recs int64[n,4] = frame, minus, start, len)."""
    a = np.frombuffer(seq, dtype=np.uint8)
    nb = ctypes.c_int64(0)
    n = lib().mo_sixframe(_p(a), ctypes.c_int64(a.size), ctypes.c_int64(min_aa), None, None, ctypes.byref(nb))
    aa = np.empty(max(nb.value, 1), dtype=np.uint8)
    rec = np.zeros((max(n, 1), 4), dtype=np.int64)
    lib().mo_sixframe(_p(a), ctypes.c_int64(a.size), ctypes.c_int64(min_aa), _p(aa), _p(rec), ctypes.byref(nb))
    return aa[:nb.value].tobytes(), rec[:n]
