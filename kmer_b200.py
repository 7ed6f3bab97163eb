#!/usr/bin/env python
"""Drop-in for `python kmer_numba.py -m -i input.fasta -k 27 > result.tab` on a B200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pangenome_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    main()
