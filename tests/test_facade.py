"""The C++ facade headers (ternary_birch_b200/host): the reference's class API over the C ABI."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def facade_binary(tmp_path_factory):
    import __graft_entry__ as ge
    ge.build()
    out = tmp_path_factory.mktemp("facade") / "facade_test"
    lib = ROOT / "ternary_birch_b200" / "libtbirch.so"
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-o", str(out), str(ROOT / "tests" / "cpp" / "facade_test.cpp"),
           str(lib), f"-Wl,-rpath,{lib.parent}"]
    subprocess.run(cmd, check=True, cwd=str(ROOT))
    return out


@pytest.fixture(scope="module")
def facade_binary_gmpxx(tmp_path_factory):
    """The same program with R = mpz_class, so host/birch.h's IntegerTraits<mpz_class> (the
    instantiation the Cython binding uses for precise=True) is compiled and run.  The image has
    libgmp.so.10 but no GMP headers: oracle/shim supplies the declarations (test infrastructure)."""
    out = tmp_path_factory.mktemp("facade") / "facade_test_gmpxx"
    lib = ROOT / "ternary_birch_b200" / "libtbirch.so"
    gmp = "/usr/lib/x86_64-linux-gnu/libgmp.so.10"
    if not Path(gmp).exists():
        pytest.skip("no libgmp.so.10 on this machine")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-DTB_FACADE_GMPXX", "-I", str(ROOT / "oracle" / "shim"),
           "-o", str(out), str(ROOT / "tests" / "cpp" / "facade_test.cpp"), str(lib), gmp, f"-Wl,-rpath,{lib.parent}"]
    subprocess.run(cmd, check=True, cwd=str(ROOT))
    return out


def test_facade_compiles_with_mpz_class(facade_binary_gmpxx):
    assert facade_binary_gmpxx.exists()


def test_facade_compiles_against_the_cabi(facade_binary):
    """Same class templates / method names the pyx declares (pyx:45-118) compile with -Wall -Werror."""
    assert facade_binary.exists()
    for name in ("birch.h", "QuadForm.h", "Isometry.h", "Eigenvector.h", "Genus.h", "IsometrySequence.h"):
        assert (ROOT / "ternary_birch_b200" / "host" / name).exists()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["stand_in", "mpz_class"])
def test_facade_matches_python_path(request, which, tables, tmp_path):
    facade_binary = request.getfixturevalue("facade_binary" if which == "stand_in" else "facade_binary_gmpxx")
    from ternary_birch_b200 import Genus, IsometrySequence
    t = tables("85085")
    stream = tmp_path / "g.tbg"
    t.to_stream().tofile(stream)
    out = tmp_path / "out.bin"
    subprocess.run([str(facade_binary), str(stream), "19", str(out)], check=True)
    a = np.fromfile(out, dtype=np.int64)
    pos = 0
    for precise in (False, True):
        g = Genus(t, precise=precise)
        want = g.hecke_matrix_sparse(19)
        C = int(a[pos]); pos += 1
        assert C == 32
        for _ in range(C):
            cond, dim, nnz = (int(x) for x in a[pos:pos + 3]); pos += 3
            data = a[pos:pos + nnz]; pos += nnz
            indices = a[pos:pos + nnz]; pos += nnz
            indptr = a[pos:pos + dim + 1]; pos += dim + 1
            assert a[pos] == 1, "dense != sparse in the facade"; pos += 1
            assert np.array_equal(data, want[cond][0]) and np.array_equal(indices, want[cond][1])
            assert np.array_equal(indptr, want[cond][2])
        seq = IsometrySequence(g, 19)
        rec = seq.read(8)
        for i in range(8):
            got = a[pos:pos + 6]; pos += 6
            assert got.tolist() == [rec[i][0], rec[i][4], rec[i][8], rec[i][9], rec[i][10], rec[i][11]]
        seq.close()
        assert a[pos] == 1, "bad-prime exception text"; pos += 1
        g.close()
    assert a[pos] == 130
    assert a[pos + 1] == 130 and a[pos + 2] == 1, "Genus(q, symbols, seed) built through the facade differs from the reference's table"
