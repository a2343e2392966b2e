"""CPU: the plain-C oracle port (and, where it exists, the compiled reference)
against the committed golden vectors generated from the unmodified reference."""
import json

import numpy as np
import pytest

from conftest import csr_digest, primes_upto
from oracle.pyoracle import RefBinary, PortError


def _digests(mats):
    return [csr_digest(m["conductor"], m["dim"], m["data"], m["indices"], m["indptr"]) for m in mats]


def test_c1_every_entry(port_oracle, hecke_small):
    """BirchGenus(11): T_p for all p < 100, all conductors, both integer modes (config 1)."""
    port = port_oracle("11")
    for p in primes_upto(100):
        if p == 11:
            continue
        for precise in (False, True):
            mats = port.hecke(p, precise)
            for k, m in enumerate(mats):
                tag = f"L11_p{p}_{'z' if precise else 'z64'}_k{k}"
                assert np.array_equal(m["data"], hecke_small[tag + "_data"]), tag
                assert np.array_equal(m["indices"], hecke_small[tag + "_indices"]), tag
                assert np.array_equal(m["indptr"], hecke_small[tag + "_indptr"]), tag
            # every row of the conductor-1 matrix sums to p + 1
            m = mats[0]
            sums = np.add.reduceat(m["data"], m["indptr"][:-1]) if m["data"].size else []
            assert all(int(s) == p + 1 for s in sums)


def test_c1_known_matrices(port_oracle):
    """Known answers for level 11: T_2, T_3, T_5, T_7, T_13."""
    port = port_oracle("11")
    want = {2: [[1, 2], [3, 0]], 3: [[2, 2], [3, 1]], 5: [[4, 2], [3, 3]], 7: [[4, 4], [6, 2]], 13: [[10, 4], [6, 8]]}
    for p, mat in want.items():
        got = port.hecke(p, False, sparse=False)[0]["matrix"]
        assert got.tolist() == mat


def test_even_level_seeded_order(port_oracle, hecke_small):
    port = port_oracle("210_s7")
    for p in (11, 13, 59):
        for k, m in enumerate(port.hecke(p, False)):
            tag = f"L210_p{p}_z64_k{k}"
            assert np.array_equal(m["data"], hecke_small[tag + "_data"]), tag
            assert np.array_equal(m["indices"], hecke_small[tag + "_indices"]), tag
            assert np.array_equal(m["indptr"], hecke_small[tag + "_indptr"]), tag


@pytest.mark.parametrize("key", [
    "L85085_p2_z64", "L85085_p2_z", "L85085_p3_z64", "L85085_p19_z", "L85085_p101_z64", "L85085_p101_z",
    "L85085_p997_z64", "L1062347_p2_z64", "L1062347_p3_z", "L1062347_p101_z64", "L1062347_p101_z",
    "L1009091_p2_z64", "L1009091_p107_z64",
])
def test_hecke_digests(port_oracle, hecke_digests, key):
    level, p, mode = key.split("_")
    port = port_oracle(level[1:])
    got = _digests(port.hecke(int(p[1:]), mode == "z"))
    assert got == hecke_digests[key]


def test_neighbor_records_and_isometry_sequences(port_oracle, records_small):
    for key, want in records_small.items():
        level, p, mode, kind = key.split("_")
        name = {"L11": "11", "L210": "210_s7", "L85085": "85085"}[level]
        port = port_oracle(name)
        if kind == "rec":
            got = port.neighbor_records(int(p[1:]), mode == "z")
        else:
            got = port.isometry_sequence(int(p[1:]), mode == "z")
        assert np.array_equal(got, want), key


def test_isometry_sum_is_hecke_row(port_oracle, records_small):
    """README identity: tallying isometry_sequence(31) by (src, dst) gives T_31 (level 11)."""
    port = port_oracle("11")
    iso = records_small["L11_p31_z64_iso"]
    T = np.zeros((2, 2), dtype=np.int64)
    for rec in iso:
        T[rec[10], rec[11]] += 1
    assert np.array_equal(T, port.hecke(31, False, sparse=False)[0]["matrix"])


def _lift(table, vec, conductor):
    k = int(np.nonzero(table.conductors == conductor)[0][0])
    lut = table.lut[k]
    full = np.zeros(table.N, dtype=np.int32)
    full[lut != -1] = np.asarray(vec, dtype=np.int32)[lut[lut != -1]]
    return full, k


def test_eigenvalues_level_11(port_oracle, tables, golden_dir):
    """a_p of the elliptic curve 11a, including primes beyond 2^16 (precise mode)."""
    g = json.load(open(golden_dir / "eigen_11.json"))
    t = tables("11")
    port = port_oracle("11")
    lifted = [_lift(t, v["vector"], v["conductor"]) for v in g["vectors"]]
    vecs = np.stack([l[0] for l in lifted])
    cond = np.array([l[1] for l in lifted])
    rep = np.array([int(np.nonzero(l[0])[0][0]) for l in lifted])
    for p, want in g["aps_z"].items():
        if int(p) > 100000:
            continue
        got = port.eigenvalues(int(p), True, vecs, cond, rep)
        assert got.tolist() == want, p


def test_eigenvalues_c3(port_oracle, tables, golden_dir):
    """Config 3: eigenvalues of the rational eigenvectors of level 5*7*11*13*17 for p <= 1000."""
    g = json.load(open(golden_dir / "eigen_85085.json"))
    t = tables("85085")
    port = port_oracle("85085")
    lifted = [_lift(t, v["vector"], v["conductor"]) for v in g["vectors"]]
    vecs = np.stack([l[0] for l in lifted])
    cond = np.array([l[1] for l in lifted])
    rep = np.array([int(np.nonzero(l[0])[0][0]) for l in lifted])
    for p in g["primes"]:
        assert port.eigenvalues(p, False, vecs, cond, rep).tolist() == g["aps_z64"][str(p)], p
    for p in g["primes"][::7]:
        assert port.eigenvalues(p, True, vecs, cond, rep).tolist() == g["aps_z"][str(p)], p
    assert port.eigenvalues(65537, True, vecs, cond, rep).tolist() == g["aps_z_big"]["65537"]


def test_bad_prime(port_oracle):
    with pytest.raises(PortError) as e:
        port_oracle("11").hecke(11, False)
    assert e.value.code == 3


@pytest.mark.skipif(not RefBinary.available(), reason="oracle/_ref not built here")
def test_reference_binary_reproduces_goldens(hecke_digests, hecke_small):
    ref = RefBinary(85085, 1)
    assert _digests(ref.hecke(101, False, True)) == hecke_digests["L85085_p101_z64"]
    assert _digests(ref.hecke(19, True, True)) == hecke_digests["L85085_p19_z"]
    ref = RefBinary(11, 1)
    m = ref.hecke(97, True, True)[0]
    assert np.array_equal(m["data"], hecke_small["L11_p97_z_k0_data"])
    # dense == sparse in the reference itself
    d = ref.hecke(97, False, False)[0]["matrix"]
    s = np.zeros_like(d)
    for r in range(2):
        s[r, m["indices"][m["indptr"][r]:m["indptr"][r + 1]]] = m["data"][m["indptr"][r]:m["indptr"][r + 1]]
    assert np.array_equal(d, s)
