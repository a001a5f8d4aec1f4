"""seqkit_b200 -- B200 (sm_100a) implementation of annalam/seqkit's per-read FASTQ batch path.

The product is libseqkit_b200.so (C ABI in include/seqkit_b200.h) and the `fasta` host binary.
This package is the thin Python binding used by the tests and bench.py; it fails loudly when the
CUDA library is missing and never falls back to a CPU implementation.
"""
from ._lib import lib, LIB_PATH, SkError  # noqa: F401
from .engine import Engine, DemuxResult  # noqa: F401
