#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel: joins the SASS page of an ncu report with the line
table nvdisasm prints for the cubin (both list the kernel's instructions in address order).

    python profiles/line_profile.py gpurun_out/prof.ncu-rep k2a_partition_c pangenome_b200/libpgdbg.so [top_n]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
            hdr = None
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and cur is not None and len(r) >= 8:
            cur["rows"].append(dict(zip(hdr, r)))
    return blocks[0] if blocks else None


def line_table(so, kernel_mangled_part):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    res = []
    for cubin in glob.glob(os.path.join(d, "*.cubin")):
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        sect, line, fname = None, None, None
        for ln in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                sect = m.group(1)
                res.append((sect, []))
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                fname, line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            if sect and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                res[-1][1].append((fname, line, ln.split("*/", 1)[1].strip()))
    return [(s, ins) for s, ins in res if kernel_mangled_part in s]


def main():
    rep, kernel, so = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    blk = sass_rows(rep, kernel)
    rows = blk["rows"]
    cands = line_table(so, re.sub(r"[^A-Za-z0-9_]", "", kernel))
    # choose the section whose instruction count matches
    cands = [c for c in cands if len(c[1]) == len(rows)] or cands
    sect, ins = cands[0]
    print("# kernel %s\n# section %s: %d SASS instructions (ncu lists %d)" % (blk["name"], sect, len(ins), len(rows)))
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0, 0.0])
    tot_i = tot_s = tot_w = 0.0
    for (f, l, txt), r in zip(ins, rows):
        n = float(r.get("Instructions Executed") or 0)
        s = float(r.get("# Samples") or 0)
        wv = float(r.get("L1 Wavefronts Shared") or 0)
        a = agg[(f, l)]
        a[0] += n; a[1] += s; a[2] += 1; a[3] += wv
        tot_i += n; tot_s += s; tot_w += wv
    print("# total warp instructions %.0f, stall samples %.0f, shared-memory wavefronts %.0f" % (tot_i, tot_s, tot_w))
    key = 3 if (len(sys.argv) > 5 and sys.argv[5] == "smem") else 0
    print("# inst%%  samples%%  smem_wavefronts%%  sass  file:line")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        print("%6.2f  %6.2f  %6.2f  %4d  %s:%s" % (100 * a[0] / max(tot_i, 1), 100 * a[1] / max(tot_s, 1), 100 * a[3] / max(tot_w, 1), a[2], f, l))


if __name__ == "__main__":
    main()
