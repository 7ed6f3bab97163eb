#!/usr/bin/env python
"""Build ``oracle/_ref/``: a runnable, deterministic copy of the reference.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this.

The reference (``/root/reference/kmer_numba.py``, one 2155-line numba script)
does not run on this image as shipped.  This recipe reads the
reference source *where it lies*, applies five semantics-preserving textual
patches in memory and writes the result to ``oracle/_ref/ref_patched.py``
(git-ignored: a derived artefact, like an object file; reference sources are
never committed).  It also writes the two stand-ins the script needs:

* ``oracle/_ref/Bio/__init__.py`` - the reference imports ``Bio.SeqIO`` at
  module level (kmer_numba.py:7) but never uses it on the jit path;
* ``oracle/_ref/bin/mcl`` - the reference shells out to the un-vendored
  third-party binary ``mcl`` (kmer_numba.py:1911).  The stand-in writes the
  undirected connected components of the ``.xyz`` graph, the definition the
  reference's own helper ``other/test_net.py:3-12`` uses.  Cluster order:
  decreasing size, ties by smallest ``(code, v5)`` node.  PARITY UNPINNED for
  the real MCL arithmetic.

Patches:
  1. ``def __delitem__`` -> ``def delitem_``   (jitclass rejects the dunder)
  2. ``def __iter__``    -> ``def iter_``      (same)
  3. ``nb.njit(inline='always')`` -> ``nb.njit``  (numba 0.65 IR bug)
  4. ``try: rdbg_edge[k12] += 1 / except: rdbg_edge[k12] = 1`` ->
     ``if k12 in rdbg_edge`` test (typed-Dict KeyError is no longer caught)
  5. determinism: ``np.empty`` -> ``np.zeros`` for ``keys``/``values`` in
     ``oakht.__init__`` (:350-351) and ``oakht.resize`` (:439-440) - the raw
     reference reads uninitialised memory (quirk Q10).
"""
import os
import stat
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PG_REFERENCE", "/root/reference/kmer_numba.py")
OUT = os.path.join(HERE, "_ref")

MCL_STANDIN = r'''#!/usr/bin/env python3
"""Stand-in for the third-party `mcl` binary: connected components of an
--abc edge list (one tab-separated line of node names per component)."""
import sys


def main(argv):
    src = argv[1]
    out = argv[argv.index('-o') + 1]
    parent = {}

    def find(x):
        r = x
        while parent[r] != r:
            r = parent[r]
        while parent[x] != r:
            parent[x], x = r, parent[x]
        return r

    with open(src) as f:
        for line in f:
            cols = line.rstrip('\n').split('\t')
            if len(cols) < 2:
                continue
            a, b = cols[0], cols[1]
            for n in (a, b):
                if n not in parent:
                    parent[n] = n
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[ra] = rb
    comps = {}
    for n in parent:
        comps.setdefault(find(n), []).append(n)

    def nkey(n):
        c, v = n.split('_')[:2]
        return (int(c), int(v))

    groups = [sorted(g, key=nkey) for g in comps.values()]
    groups.sort(key=lambda g: (-len(g), nkey(g[0])))
    with open(out, 'w') as f:
        for g in groups:
            f.write('\t'.join(g) + '\n')


if __name__ == '__main__':
    main(sys.argv)
'''


def patch_source(src):
    n = {}

    def sub(tag, old, new, expect):
        nonlocal src
        c = src.count(old)
        if c != expect:
            raise SystemExit("make_ref: patch %s expected %d sites, found %d" % (tag, expect, c))
        src = src.replace(old, new)
        n[tag] = c

    sub("delitem", "def __delitem__", "def delitem_", 2)
    sub("iter", "def __iter__", "def iter_", 2)
    sub("inline", "nb.njit(inline='always')", "nb.njit", 4)
    sub("trydict",
        "                    try:\n"
        "                        rdbg_edge[k12] += 1\n"
        "                    except:\n"
        "                        rdbg_edge[k12] = 1\n",
        "                    if k12 in rdbg_edge:\n"
        "                        rdbg_edge[k12] += 1\n"
        "                    else:\n"
        "                        rdbg_edge[k12] = 1\n", 1)
    sub("zeros_init_k", "        self.keys = np.empty(N * ksize, dtype=ktype)",
        "        self.keys = np.zeros(N * ksize, dtype=ktype)", 1)
    sub("zeros_init_v", "        self.values = np.empty(N * vsize, dtype=vtype)",
        "        self.values = np.zeros(N * vsize, dtype=vtype)", 1)
    # resize() and resize_disk-free copy: only resize() allocates (two sites: :439-440)
    sub("zeros_resize_k", "        keys = np.empty(M * ks, dtype=keys_old.dtype)",
        "        keys = np.zeros(M * ks, dtype=keys_old.dtype)", 1)
    sub("zeros_resize_v", "        values = np.empty(M, dtype=values_old.dtype)",
        "        values = np.zeros(M, dtype=values_old.dtype)", 1)
    return src, n


def build(verbose=True):
    if not os.path.isfile(REF_SRC):
        if verbose:
            print("make_ref: %s not present; nothing built" % REF_SRC)
        return False
    with open(REF_SRC) as f:
        src = f.read()
    patched, counts = patch_source(src)
    os.makedirs(os.path.join(OUT, "Bio"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "bin"), exist_ok=True)
    with open(os.path.join(OUT, "ref_patched.py"), "w") as f:
        f.write(patched)
    with open(os.path.join(OUT, "Bio", "__init__.py"), "w") as f:
        f.write("SeqIO = None\n")
    mcl = os.path.join(OUT, "bin", "mcl")
    with open(mcl, "w") as f:
        f.write(MCL_STANDIN)
    os.chmod(mcl, os.stat(mcl).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)
    if verbose:
        print("make_ref: wrote %s (patch sites: %s)" % (OUT, counts))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
