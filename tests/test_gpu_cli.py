"""The CLI end to end on the GPU: same table as the reference printed for test.fsa."""
import io
import os
import shutil

import pytest

from conftest import load_small_cases

pytestmark = pytest.mark.gpu


def test_cli_test_fsa(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pangenome_b200 import cli
    case = {c["name"]: c for c in load_small_cases()}["test_fsa_k5"]
    fa = tmp_path / "test.fsa"
    fa.write_bytes(case["input_latin1"].encode("latin-1"))
    out = io.StringIO()
    cli.entry_point(["prog", "-m", "-i", str(fa), "-k5", "-n", "5e8"], out=out)
    lines = out.getvalue().split("\n")
    rows = [l for l in lines if l and not l.startswith("#")]
    assert rows == ["1\t0\t39\t+\t0", "2\t0\t7\t+\t1"]
    # the banner sequence of kmer_numba.py's entry_point (:2106-2144), timing figures aside
    banners = [l if not l.startswith("# finished in") else "# finished" for l in lines if l.startswith("# ")]
    assert banners == ["# build the dBG", "# finished", "# save dBG to disk", "# finished", "# load dBG from disk", "# finished",
                       "# build the reduced dBG", "# finished", "# find fr", "# finished"]
    assert os.path.isfile(str(fa) + "_db.npz")                # like upstream, the dBG is saved beside the input
    assert (tmp_path / "test.fsa_rdbg_weight.xyz").read_text().split("\n")[:-1] == case["xyz"]
    assert sorted((tmp_path / "test.fsa_rdbg_weight.xyz.mcl").read_text().split("\n")[:-1]) == sorted(case["mcl"])
    # second run: the cluster file exists -> "# the mcl has been ran", same rows
    out2 = io.StringIO()
    cli.entry_point(["prog", "-i", str(fa), "-k", "5"], out=out2)
    assert "# the mcl has been ran" in out2.getvalue()
    assert [l for l in out2.getvalue().split("\n") if l and not l.startswith("#")] == rows
    # k = 27: the reference crashes on the empty edge set (F9); defined behaviour: no rows, exit 0
    os.remove(tmp_path / "test.fsa_rdbg_weight.xyz.mcl")
    out3 = io.StringIO()
    cli.entry_point(["prog", "-i", str(fa), "-k", "27"], out=out3)
    assert [l for l in out3.getvalue().split("\n") if l and not l.startswith("#")] == []


def test_cli_dump_db_and_reload(tmp_path):
    """<input>_db.npz is written in the reference's layout (default, like upstream); -d starts from it; --no-dump-db skips it."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from pangenome_b200 import cli
    case = {c["name"]: c for c in load_small_cases()}["nasty_k11_lf"]
    fa = tmp_path / "n.fa"
    fa.write_bytes(case["input_latin1"].encode("latin-1"))
    out = io.StringIO()
    cli.entry_point(["prog", "-i", str(fa), "-k", "11", "--no-mcl-file"], out=out)
    rows = [l for l in out.getvalue().split("\n") if l and not l.startswith("#")]
    assert rows == ["%s\t%d\t%d\t%s\t%d" % tuple(r) for r in case["rows"]]
    z = np.load(str(fa) + "_db.npz")
    assert int(z["parameters"][2]) == len(case["dbg"])
    out2 = io.StringIO()
    cli.entry_point(["prog", "-i", str(fa), "-k", "11", "-d", str(fa) + "_db.npz", "--no-mcl-file"], out=out2)
    assert [l for l in out2.getvalue().split("\n") if l and "\t" in l] == rows
    os.remove(str(fa) + "_db.npz")
    out3 = io.StringIO()
    cli.entry_point(["prog", "-i", str(fa), "-k", "11", "--no-dump-db", "--no-mcl-file"], out=out3)
    assert not os.path.exists(str(fa) + "_db.npz")
    assert [l for l in out3.getvalue().split("\n") if l and "\t" in l] == rows


def test_cli_stage1_runs_the_benchmarked_kernels(tmp_path):
    """The drop-in's stage 1 is the build bench.py times: its kernel list holds k2a_partition and k3s_region_build (the
    streaming build into shared-memory table regions), not the fused single-launch kernel."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from torch.profiler import ProfilerActivity, profile
    from pangenome_b200 import stages
    from pangenome_b200.synth import pangenome
    fa = tmp_path / "p.fa"
    fa.write_bytes(pangenome(3, 60_000, seed=9))
    try:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            stages.seq2rdbg(str(fa), 27, 5, 2 ** 63, brkpt="", chunk=2 ** 33, rc=True)
            torch.cuda.synchronize()
        names = {e.key for e in prof.key_averages()}
    except Exception as e:          # CUPTI unavailable on the box
        pytest.skip("kernel tracing unavailable: %r" % (e,))
    if not names:
        pytest.skip("kernel tracing returned no events")
    assert any("k2a_partition" in n for n in names), sorted(names)
    assert any("k3s_region_build" in n for n in names), sorted(names)
    assert any("k1x_pack" in n for n in names), sorted(names)
    assert not any("k2_kmer_insert" in n for n in names), sorted(names)
