"""pangenome_b200 - B200-native de Bruijn-graph hot path of Rinoahu/pangenome.

Host-side mirror of the reference's stage functions (``seq2rdbg``, ``dbg2rdbg``,
``seq2graph``; kmer_numba.py:1234, 1313, 1853) over hand-written sm_100a CUDA
kernels behind the C-ABI of ``include/pgdbg.h``.  There is no CPU fallback: the
stage functions raise if ``libpgdbg.so`` is missing or no CUDA device exists.
"""
__version__ = "0.1.0"
