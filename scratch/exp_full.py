import sys, time, torch, numpy as np, io, os, tempfile
sys.path.insert(0, '.')
from pangenome_b200 import engine, graph, stages, cli
from pangenome_b200.synth import pangenome
data = pangenome(10, 5_000_000)
def T(name, f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); print("%-28s %8.2f ms" % (name, 1e3 * (time.perf_counter() - t)), flush=True); return r
for it in range(2):
    print("--- iteration", it)
    d = T("H2D", lambda: engine.to_device_bytes(data))
    p = T("K1 pack (+index D2H)", lambda: engine.PackedSeqs(d))
    t, n_rec, _ = T("dBG two-phase", lambda: engine.build_dbg_partitioned(p, 27))
    rd = T("K4 rdBG select", lambda: t.select_rdbg())
    print("   rdBG members", rd.n_members, "slots", rd.n_slots_used, "cap", rd.capacity)
    h = T("K5 path hits", lambda: graph.path_hits(p, rd, n_rec, 0))
    print("   hits", h.n)
    g = T("graph alloc+clear", lambda: graph.RdbgGraph(h.n, p.pk2.device))
    T("K6 add_hits", lambda: g.add_hits(h, 1))
    e = T("edges export+sort", lambda: g.edges(rd))
    print("   edges", e[0].size)
    nodes = T("K7 components+host order", lambda: g.components(rd))
    print("   nodes", nodes[0].size, "components", len(set(nodes[3].tolist())))
    T("set_labels", lambda: g.set_labels(nodes[0], nodes[3]))
    rows = T("K8 regions", lambda: g.regions(h, p, 27))
    print("   rows", rows[0].size)
    res = graph.GraphResult(); res.edges = e; res.nodes = nodes
    lines = T("xyz text (host)", lambda: res.xyz_lines())
    ml = T("mcl text (host)", lambda: res.mcl_lines())
tmp = tempfile.mkdtemp()
fa = os.path.join(tmp, "cfg2.fa"); open(fa, "wb").write(data)
t0 = time.perf_counter(); out = io.StringIO(); cli.entry_point(["prog", "-i", fa, "-k", "27"], out=out); print("CLI total %.1f ms" % (1e3 * (time.perf_counter() - t0)))
print(out.getvalue()[:600])
