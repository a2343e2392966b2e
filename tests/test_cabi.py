"""CPU: the C-ABI library loads and exports every symbol include/tbirch.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import has_gpu

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from ternary_birch_b200 import cabi
    return cabi


def declared_functions():
    text = (ROOT / "include" / "tbirch.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_header(built):
    L = built.lib()
    names = declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"libtbirch.so does not export {name}"
    assert sorted(built.EXPORTS) == names
    assert L.tb_abi_version() == 2


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built, tables):
    from ternary_birch_b200 import Genus
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Genus(tables("11"))
    peak = ctypes.c_double()
    assert built.lib().tb_int_peak(ctypes.byref(peak), None, None) == built.TB_ERR_RUNTIME


def test_product_does_not_import_oracle():
    """The product package and bench's own arm never reach into oracle/."""
    for path in (ROOT / "ternary_birch_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            text = path.read_text()
            assert "pyoracle" not in text and "tb_oracle" not in text and "oracle/" not in text, path


def test_table_roundtrip(tables, tmp_path):
    from ternary_birch_b200 import load_genus_table, save_genus_table
    t = tables("85085")
    assert (t.N, t.K) == (130, 5)
    assert t.level == 85085
    assert t.q[0].tolist() == [1, 130, 172, -65, -1, 0]
    for name in ("a.npz", "a.tbg"):
        save_genus_table(t, tmp_path / name)
        u = load_genus_table(tmp_path / name)
        assert np.array_equal(u.to_stream(), t.to_stream())
    assert tables("1062347").q[0].tolist() == [1, 187, 1467, -187, 0, 0]
    assert tables("1009091").N == 10316


def test_host_widening_pool_without_gpu():
    """The host half of the narrow read-back (uint16 / int8 -> int32 with the library's thread pool)
    needs no GPU: the probe widens a buffer, checks the values and reports its throughput."""
    import ctypes
    from ternary_birch_b200 import cabi
    L = cabi.lib()
    gbps, threads, avx2 = ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
    for entries in (1, 7, 1 << 18, (1 << 20) + 13):
        assert L.tb_host_widen_probe(entries, ctypes.byref(gbps), ctypes.byref(threads), ctypes.byref(avx2)) == 0
        assert gbps.value > 0 and threads.value >= 2 and avx2.value in (0, 1, 2)
    assert L.tb_host_widen_probe(0, None, None, None) != 0


def test_peak_microbenchmark_counts_sass_instructions():
    """The roofline denominator counts operations per SASS instruction: every step of int_peak_kernel is one
    PTX instruction in an asm statement, and the unrolled loop body must hold exactly that many instructions
    of the measured kind (16 steps x 8, 16 or 24), whatever ptxas does with them (it splits plain adds
    between IADD3 on the alu pipe and IMAD.IADD on the fma pipe)."""
    import collections
    import re
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    sass = subprocess.run(["cuobjdump", "-sass", str(Path(__file__).resolve().parents[1] / "ternary_birch_b200" / "libtbirch.so")], capture_output=True, text=True, check=True).stdout
    counts, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            counts[name][m.group(1)] += 1
    want = {0: (("IMAD",), 128), 1: (("LOP3",), 128), 2: (("IADD3", "IMAD"), 128), 3: (("IADD3", "IMAD"), 256),
            4: (("IMAD", "LOP3"), 256), 5: (("DFMA",), 128), 6: (("DFMA", "IMAD", "LOP3"), 384)}
    for kind, (ops, n) in want.items():
        c = next(v for k, v in counts.items() if f"int_peak_kernelILi{kind}E" in k)
        got = sum(c[o] for o in ops)
        assert n <= got <= n + 64, (kind, dict(c))      # the body, plus a few set-up instructions of the same kind
