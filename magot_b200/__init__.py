"""magot_b200 -- B200-native implementation of MAGOT's annotation-driven sequence path.

    from magot_b200 import genome            # drop-in for the reference's `genome` module
    from magot_b200 import genome_tools      # drop-in for the `genome_tools` entry points

The CUDA library (magot_b200/libmagot_b200.so, built by `__graft_entry__.build()`) is required;
there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library has not been built)
from . import genome  # noqa: F401
from .genome import (Sequence, GenomeSequence, Genome, AnnotationSet, ParentAnnotation,  # noqa: F401
                     BaseAnnotation, read_gff, write_gff, position_dic)

__version__ = "0.1.0"
