"""magot_b200 -- B200-native implementation of MAGOT's annotation-driven sequence path.

    from magot_b200 import genome            # drop-in for the reference's `genome` module
    from magot_b200 import genome_tools      # drop-in for the `genome_tools` entry points

The CUDA library (magot_b200/libmagot_b200.so, built by `__graft_entry__.build()`) is required by every
module that computes (`_lib`, `engine`, `genome`, `genome_tools`, `orfs`, `flatten`): importing any of them
fails loudly when it has not been built; there is no CPU fallback.  Submodules are imported on first use, so
that the pure-numpy helpers (`synth`, `tables`, `py2dict`: workload generators and table containers, used by
bench.py's reference arm) do not map the CUDA library into a process that must not run it.
"""
import importlib

__version__ = "0.2.0"

_SUBMODULES = ("_lib", "engine", "genome", "genome_tools", "orfs", "flatten", "synth", "tables", "py2dict", "gffnative")
_FROM_GENOME = ("Sequence", "GenomeSequence", "Genome", "AnnotationSet", "ParentAnnotation", "BaseAnnotation",
                "read_gff", "write_gff", "position_dic")


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    if name in _FROM_GENOME:
        return getattr(importlib.import_module(".genome", __name__), name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def __dir__():
    return sorted(list(globals()) + list(_SUBMODULES) + list(_FROM_GENOME))
