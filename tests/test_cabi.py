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
