"""Splitting ONE input file across the ranks of a multi-GPU run: rank r takes the byte range
[cut_r, cut_{r+1}) where cut_r is the first record start ('>' at the beginning of a line) at or after the nominal
offset r * size / world.  Ranges are record-aligned, so every rank packs whole records and the k-mers of a record never
straddle ranks; no communication is needed to agree on the cuts (every rank can compute all of them from the file).
A record larger than size / world is not split (the ranks whose nominal offsets fall inside it get empty ranges).

Host logic only - no CUDA here."""
import mmap
import os


def cut_points_from_starts(starts, size, world):
    """starts: ascending byte offsets of the record starts.  Returns world + 1 cut offsets."""
    import bisect
    cuts = [0]
    for r in range(1, world):
        nominal = r * size // world
        j = bisect.bisect_left(starts, nominal)
        cuts.append(starts[j] if j < len(starts) else size)
    cuts.append(size)
    # the bytes before the first header belong to rank 0 (they are dropped by the reader anyway)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


def _first_record_start(buf, size, nominal):
    """First offset >= nominal that holds a '>' at a line start."""
    if nominal <= 0:
        return 0
    j = buf.find(b"\n>", nominal - 1)
    return size if j < 0 else j + 1


def cut_points(buf, world):
    """Cut offsets for a bytes-like / mmap object."""
    size = len(buf)
    cuts = [0] + [_first_record_start(buf, size, r * size // world) for r in range(1, world)] + [size]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


def read_rank_range(path, world, rank):
    """(bytes of this rank's record-aligned range, (begin, end), file size).  Only the rank's own pages are read
    (plus the few a cut search touches)."""
    size = os.path.getsize(path)
    if size == 0:
        return b"", (0, 0), 0
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            cuts = cut_points(mm, world)
            a, b = cuts[rank], cuts[rank + 1]
            data = mm[a:b]
        finally:
            mm.close()
    return data, (a, b), size
