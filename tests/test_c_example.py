"""examples/count_kmers.c: the C-ABI from plain C (gcc, no Python, no torch).
CPU: include/pgdbg.h is valid C and the example links against libpgdbg.so.
GPU: its output for golden inputs equals the oracle's table checksum."""
import os
import subprocess

import pytest

import oracle
from conftest import load_small_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from pangenome_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("libpgdbg.so not built")
    out = str(tmp_path_factory.mktemp("cex") / "count_kmers")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + CUDA + "/include",
                           os.path.join(ROOT, "examples", "count_kmers.c"), "-o", out, "-L" + libdir, "-lpgdbg",
                           "-L" + CUDA + "/lib64", "-lcudart", "-Wl,-rpath," + libdir])
    return out


def test_example_compiles_as_c(exe):
    assert os.access(exe, os.X_OK)
    r = subprocess.run([exe], capture_output=True, text=True)      # no arguments: usage, no CUDA call
    assert r.returncode == 1 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_fsa_k27", "nasty_k11_lf", "nasty_k27_crlf", "nasty_k5_nofinal"])
def test_example_matches_oracle(exe, name, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    case = [c for c in load_small_cases() if c["name"] == name][0]
    data = case["input_latin1"].encode("latin-1")
    fa = tmp_path / "in.fa"
    fa.write_bytes(data)
    r = subprocess.run([exe, str(fa), str(case["k"])], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    ref = oracle.run(data, case["k"], stages=1)
    n, s, x = oracle.table_checksum(*ref["dbg"])
    first = r.stdout.split("\n")[0].split()
    assert first[-3:] == [str(n), str(s), str(x)], r.stdout
    assert int(first[first.index("records") + 1]) == len(ref["seq_off"]) - 1
    assert "same checksum" in r.stdout
