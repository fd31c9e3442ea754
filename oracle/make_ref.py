#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- builds `oracle/_ref/`, an importable copy of the reference.

The reference (MAGOT: genome.py, genome_tools.py, magot_smallfuncs.py,
genome_tools_config.py) is Python-2.7-only source; neither this container nor
the GPU box has a Python 2 interpreter.  This script reads the four files where
they lie under /root/reference and writes a *mechanically* shimmed copy into the
git-ignored directory `oracle/_ref/` so that the reference's OWN code can be
executed under CPython 3 as (a) the validation target of the oracle restatement
(`oracle/magot_oracle.py`, `oracle/oracle.c`), (b) the generator of the golden
vectors under tests/golden/ and (c) the `--impl reference` arm of bench.py.

Nothing under oracle/_ref/ is tracked by git (it is a build product, like a
compiled reference binary would be) and no product module imports it.

The rewrite rules are purely syntactic and never touch an algorithm:
  1. `print <expr>` statements  ->  `print(<expr>)`   (joining continuation
     lines while a triple quote is open or the line ends in a backslash:
     genome.py:394-396, :612-613, :673-674; genome_tools.py:61-62)
  2. `import StringIO`          ->  `import io as StringIO`
     (genome.py:17, magot_smallfuncs.py:9)
  3. `type(output) == file`     ->  `False`            (genome.py:1078)
  4. `len(self[seqid][:-1 * window_size]) / window_jump` -> `//` (genome.py:1052: Python 2's `/` on two ints IS floor
     division; the rule only names that one expression)
  5. a prologue line `from _py2compat import open` so that text files are read
     as Python 2 reads them: bytes one-to-one (latin-1) and *no* universal
     newline translation (a lone '\r' stays inside its line).
Semantic differences that remain (all OFF the hot path): `presets=` of read_gff
(`exec` cannot rebind locals in Py3),
`ensure_file` on already-open file objects, and dict iteration order (Python 3
dicts iterate in insertion order; CPython 2.7 iterates in hash-slot order -- the
order emulator lives in oracle/py2dict.py and is applied by the callers).
"""
import os
import re
import sys

REF = os.environ.get("MAGOT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["genome.py", "genome_tools.py", "magot_smallfuncs.py", "genome_tools_config.py"]

COMPAT = '''"""py2 text-file semantics for the shimmed reference (test infrastructure)."""
import builtins as _b


def open(name, mode="r", *a, **k):
    if "b" not in mode:
        k.setdefault("encoding", "latin-1")
        k.setdefault("newline", "\\n")
    return _b.open(name, mode, *a, **k)
'''

_PRINT = re.compile(r"^(\s*)print\s+(.*)$", re.S)


def shim_source(text):
    lines = text.split("\n")
    out = []
    i = 0
    while i < len(lines):
        line = lines[i]
        m = _PRINT.match(line)
        if m and not line.lstrip().startswith("#"):
            indent, expr = m.group(1), m.group(2)
            # join continuation lines: open triple quote or trailing backslash
            while (expr.count('"""') % 2 == 1) or expr.rstrip().endswith("\\"):
                i += 1
                expr = expr + "\n" + lines[i]
            out.append(indent + "print(" + expr + ")")
        else:
            out.append(line)
        i += 1
    text = "\n".join(out)
    text = text.replace("import StringIO", "import io as StringIO")
    text = text.replace("type(output) == file", "False")
    text = text.replace("len(self[seqid][:-1 * window_size]) / window_jump", "len(self[seqid][:-1 * window_size]) // window_jump")
    return "from _py2compat import open\n" + text


def available():
    return all(os.path.isfile(os.path.join(REF, f)) for f in FILES)


def build(force=False):
    """Write oracle/_ref/*.py. Returns the directory, or None when the reference is absent."""
    if not available():
        return OUT if os.path.isfile(os.path.join(OUT, "genome.py")) else None
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "_py2compat.py"), "w") as fh:
        fh.write(COMPAT)
    for f in FILES:
        with open(os.path.join(REF, f), encoding="latin-1", newline="\n") as fh:
            src = fh.read()
        dst = os.path.join(OUT, f)
        new = shim_source(src)
        if force or not os.path.isfile(dst) or open(dst, encoding="latin-1", newline="\n").read() != new:
            with open(dst, "w", encoding="latin-1", newline="\n") as fh:
                fh.write(new)
    return OUT


def load():
    """Import the shimmed reference's `genome` module (None when unavailable)."""
    d = build()
    if d is None:
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    import importlib
    return importlib.import_module("genome")


if __name__ == "__main__":
    d = build(force=True)
    print("reference shim:", d if d else "unavailable (no /root/reference)")
