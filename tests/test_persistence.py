"""Result persistence: the sqlite eigenvector database
(reference: src/sage/build_database.py) and the per-level JSON documents
(reference: scripts/generate_eigenvalues.py).

CPU: schema, packing, oldform pruning, level plans and the JSON layout against a stub
genus.  GPU: the whole batch flow on small levels, checked against FIRST-PRINCIPLES known
answers: by modularity the rational eigenvectors at a squarefree level are the isogeny
classes of elliptic curves of that conductor, and a_p = p + 1 - #E(F_p) is counted here
directly from Cremona's minimal models."""
import json
import sqlite3

import pytest

from ternary_birch_b200 import batch
from ternary_birch_b200.database import EigenvectorDatabase


class StubGenus:
    """The slice of BirchGenus the persistence layer touches."""
    calls = []

    def __init__(self, level, ramified_primes=None, seed=None):
        StubGenus.calls.append((level, ramified_primes, seed))
        self.level_, self.seed_ = int(level), 77 if seed is None else seed
        primes = batch.prime_divisors(self.level_)
        self.ramified_primes_ = list(ramified_primes) if ramified_primes else (primes[:-1] if len(primes) % 2 == 0 else primes)
        self.eigenvectors = None
        self.resets = 0

    def level(self):
        return self.level_

    def seed(self):
        return self.seed_

    def ramified_primes(self):
        return self.ramified_primes_

    def dimensions(self):
        return {1: 3, self.level_: 2}

    def dimension(self):
        return 5

    def rational_eigenvectors(self, precise=True):
        if self.eigenvectors is None:
            q = self.level_ // self.ramified_primes_[0] if len(batch.prime_divisors(self.level_)) == 2 else 1
            self.eigenvectors = [dict(vector=[1, -2, 0], aps={2: -1}, conductor=1),
                                 dict(vector=[0, 3, 1], aps={2: 1}, conductor=1)]
            if q > 1:                                   # an oldform pair: conductors d and d * q, same a_p
                self.eigenvectors += [dict(vector=[5, 5], aps={2: 2}, conductor=self.ramified_primes_[0]),
                                      dict(vector=[1, 4], aps={2: 2}, conductor=self.level_)]
        return self.eigenvectors

    def add_eigenvector(self, vec):
        if self.eigenvectors is None:
            self.eigenvectors = []
        self.eigenvectors.append(dict(vec))

    def reset_eigenvector_manager(self):
        self.resets += 1

    def compute_eigenvalues_upto(self, upper, precise=True):
        for v in self.rational_eigenvectors():
            for p in (3, 5, 7):
                if p <= upper and self.level_ % p:
                    v["aps"][p] = v["aps"][2] * p


def test_database_schema_and_roundtrip(tmp_path):
    path = tmp_path / "eigenvectors.sqlite3"
    with EigenvectorDatabase(str(path), genus_factory=StubGenus) as db:
        g = db.get_genus_with_eigenvectors(33, precise=False)        # 3 * 11: two primes -> oldform removal
        assert [v["conductor"] for v in g.eigenvectors] == [1, 1] and g.resets == 1
        assert all("vector_id" in v for v in g.eigenvectors)
        g.eigenvectors[0]["aps"][13] = 99
        db.update_database_from_genus(g, precise=False)
        g2 = db.get_genus_with_eigenvectors(30, precise=False, remove_oldforms=True)   # three primes: nothing removed
        assert len(g2.eigenvectors) == 2 and g2.resets == 0
        db.commit()
    conn = sqlite3.connect(str(path))
    cols = {t: [r[1] for r in conn.execute(f"PRAGMA table_info({t})")] for t in ("genus", "eigenvectors", "eigenvalues")}
    assert cols == dict(genus=["level", "ramified_primes", "num_eigenvectors", "seed"],
                        eigenvectors=["id", "level", "conductor", "vector"],
                        eigenvalues=["vector_id", "p", "value"])
    assert conn.execute("SELECT level, ramified_primes, num_eigenvectors, seed FROM genus ORDER BY level").fetchall() == \
        [(30, "2,3,5", 2, 77), (33, "3", 2, 77)]
    assert conn.execute("SELECT vector FROM eigenvectors WHERE level=33 ORDER BY id").fetchall() == [("1,-2,0",), ("0,3,1",)]
    assert (1, 13, 99) in conn.execute("SELECT vector_id, p, value FROM eigenvalues").fetchall()
    conn.close()
    # a second session rebuilds the genus from the stored parameters and reloads the vectors
    StubGenus.calls.clear()
    with EigenvectorDatabase(str(path), genus_factory=StubGenus) as db:
        g = db.get_genus_with_eigenvectors(33, precise=False)
        assert StubGenus.calls == [(33, [3], 77)]
        assert [v["vector"] for v in g.eigenvectors] == [[1, -2, 0], [0, 3, 1]]
        assert g.eigenvectors[0]["aps"][13] == 99 and g.eigenvectors[0]["aps"][2] == -1
        with pytest.raises(LookupError, match="is not in"):
            db.update_database_from_genus(StubGenus(35), precise=False)
    assert EigenvectorDatabase.unpack_vector(EigenvectorDatabase.pack_vector((3, -4, 0))) == [3, -4, 0]
    # conductor 33 = 3 * 11 with the eigenvalues of the conductor-3 vector wherever both know them: an oldform pair
    assert EigenvectorDatabase.identify_duplicates(
        [dict(conductor=1, aps={2: 0}), dict(conductor=3, aps={2: 1, 5: 2}), dict(conductor=33, aps={2: 1, 7: 0}),
         dict(conductor=33, aps={2: -1})], 11) == [1, 2]


def test_level_plans_and_json(tmp_path):
    assert [batch.sturm_bound(n) for n in (11, 13, 379, 99)] == [2, 3, 64, 24]
    assert batch.is_squarefree(30, num_primes=3) and not batch.is_squarefree(12) and not batch.is_squarefree(30, num_primes=2)
    assert not batch.is_squarefree(1) and batch.squarefree_primes(1001) == [7, 11, 13] and batch.squarefree_primes(49) is None
    plan = batch.plan_level(11 * 13, seed=3)
    assert (plan.ramified, plan.unramified, plan.ramified_product, plan.unramified_product) == ([11], [13], 11, 13)
    assert plan.sturm == 2 and plan.upto == 2
    plan = batch.plan_level(2 * 3 * 5, seed=3, upto=5)
    assert (plan.ramified, plan.unramified, plan.upto) == ([2, 3, 5], [], 5)
    assert batch.plan_level(2 * 3 * 5, seed=3, upto=10 ** 6).upto == batch.sturm_bound(30)
    with pytest.raises(ValueError):
        batch.plan_level(12, seed=1)
    doc = batch.run_level(batch.plan_level(1001, seed=9, upto=7), genus_factory=StubGenus)
    path = batch.write_document(doc, tmp_path)
    assert path.endswith("1001.json")
    raw = json.loads(open(path).read())
    # the reference's document: same keys, same order (scripts/generate_eigenvalues.py:22-47)
    assert list(raw) == ["job", "dimensions_by_conductor", "dimension_total", "rational_eigenvectors_count",
                         "rational_eigenvectors"]
    assert list(raw["job"]) == ["level", "seed", "ramified_primes", "ramified_primes_product", "unramified_primes",
                                "unramified_primes_product", "sturm_bound_of_ramified_primes_product",
                                "eigenvalues_upto", "elapsed_seconds", "elapsed_seconds_genus",
                                "elapsed_seconds_eigenvectors", "elapsed_seconds_eigenvalues"]
    assert raw["job"]["elapsed_seconds"] >= raw["job"]["elapsed_seconds_eigenvalues"] >= 0
    back = batch.load_result(path)
    assert back["dimensions_by_conductor"] == {1: 3, 1001: 2} and back["rational_eigenvectors_count"] == 2
    assert back["rational_eigenvectors"][0] == dict(vector=[1, -2, 0], aps={2: -1, 3: -3, 5: -5}, conductor=1)


# Cremona minimal models [a1, a2, a3, a4, a6], one per isogeny class of the conductor
CURVES = {
    11: [[0, -1, 1, -10, -20]],
    14: [[1, 0, 1, 4, -6]],
    15: [[1, 1, 1, -10, -10]],
    26: [[1, 0, 1, -5, -8], [1, -1, 1, -3, 3]],
    30: [[1, 0, 1, 1, 2]],
    37: [[0, 0, 1, -1, 0], [0, 1, 1, -23, -50]],
}


def trace_of_frobenius(curve, p):
    a1, a2, a3, a4, a6 = curve
    count = 1                                            # the point at infinity
    for x in range(p):
        rhs = (x * x * x + a2 * x * x + a4 * x + a6) % p
        for y in range(p):
            if (y * y + a1 * x * y + a3 * y - rhs) % p == 0:
                count += 1
    return p + 1 - count


@pytest.mark.gpu
def test_batch_flow_against_elliptic_curves(tmp_path):
    """scripts/generate_eigenvalues.py's per-level job and build_database.py's flow, on real genera
    built on the GPU.  Levels with two prime factors exercise the oldform removal."""
    primes = [p for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31)]
    with EigenvectorDatabase(str(tmp_path / "e.sqlite3")) as db:
        for level, curves in CURVES.items():
            genus = db.get_genus_with_eigenvectors(level, precise=False, remove_oldforms=True)
            vecs = genus.rational_eigenvectors()
            genus.compute_eigenvalues_upto(31, precise=False)
            db.update_database_from_genus(genus, precise=False)
            good = [p for p in primes if level % p]
            got = sorted(tuple(v["aps"][p] for p in good) for v in vecs)
            want = sorted(tuple(trace_of_frobenius(c, p) for p in good) for c in curves)
            assert got == want, level
        db.commit()
        again = db.get_genus_with_eigenvectors(37, precise=False)
        assert sorted(v["aps"][2] for v in again.eigenvectors) == [-2, 0]
        assert again.compute_eigenvalues(41, precise=True) is not None
    doc = batch.run_level(batch.plan_level(37, seed=1))
    assert doc["rational_eigenvectors_count"] == 2 and doc["dimension_total"] == sum(doc["dimensions_by_conductor"].values())
    assert doc["job"]["eigenvalues_upto"] == batch.sturm_bound(37) == 7
    assert sorted(v["aps"][7] for v in doc["rational_eigenvectors"]) == [-1, -1]


# Number of isogeny classes of elliptic curves of each squarefree conductor 11 <= N <= 130 (Cremona's tables):
# by modularity, the number of rational newforms of level N.
CREMONA_CLASS_COUNTS = {
    11: 1, 13: 0, 14: 1, 15: 1, 17: 1, 19: 1, 21: 1, 22: 0, 23: 0, 26: 2, 29: 0, 30: 1, 31: 0, 33: 1, 34: 1, 35: 1, 37: 2,
    38: 2, 39: 1, 41: 0, 42: 1, 43: 1, 46: 1, 47: 0, 51: 1, 53: 1, 55: 1, 57: 3, 58: 2, 59: 0, 61: 1, 62: 1, 65: 1, 66: 3,
    67: 1, 69: 1, 70: 1, 71: 0, 73: 1, 74: 0, 77: 3, 78: 1, 79: 1, 82: 1, 83: 1, 85: 1, 86: 0, 87: 0, 89: 2, 91: 2, 93: 0,
    94: 1, 95: 0, 97: 0, 101: 1, 102: 3, 103: 0, 105: 1, 106: 4, 107: 0, 109: 1, 110: 3, 111: 0, 113: 1, 114: 3, 115: 1,
    118: 4, 119: 0, 122: 1, 123: 2, 127: 0, 129: 2, 130: 3,
}


@pytest.mark.gpu
def test_rational_newform_counts_match_cremona(tmp_path):
    """Every squarefree level from 11 to 130 through the batch flow (genus on the GPU, eigenvector search,
    oldform removal at levels with an even number of prime factors): the number of rational eigenvectors
    equals the number of isogeny classes of elliptic curves of that conductor."""
    got = {}
    with EigenvectorDatabase(str(tmp_path / "e.sqlite3")) as db:
        for level in CREMONA_CLASS_COUNTS:
            assert batch.is_squarefree(level)
            got[level] = len(db.get_genus_with_eigenvectors(level, precise=False).rational_eigenvectors())
    assert got == CREMONA_CLASS_COUNTS


@pytest.mark.gpu
def test_generate_eigenvalues_script(tmp_path):
    """The batch driver's command line (reference: scripts/generate_eigenvalues.py:169-231): one JSON per
    squarefree level of the range, loadable, with eigenvalues up to the Sturm bound of the ramified part."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = tmp_path / "json"
    r = subprocess.run([sys.executable, str(root / "scripts" / "generate_eigenvalues.py"), str(out), "--lower", "33",
                        "--upper", "40", "--seed", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    levels = sorted(int(f.stem) for f in out.glob("*.json"))
    assert levels == [33, 34, 35, 37, 38, 39]
    doc = batch.load_result(out / "37.json")
    assert doc["job"]["seed"] == 5 and doc["job"]["ramified_primes"] == [37] and doc["job"]["eigenvalues_upto"] == 7
    assert doc["rational_eigenvectors_count"] == 2
    assert sorted(tuple(v["aps"][p] for p in (2, 3, 5, 7)) for v in doc["rational_eigenvectors"]) == \
        sorted(tuple(trace_of_frobenius(c, p) for p in (2, 3, 5, 7)) for c in CURVES[37])
    doc = batch.load_result(out / "38.json")          # 2 * 19: ramified [2], the search keeps oldforms (as the reference's script does)
    assert doc["job"]["ramified_primes"] == [2] and doc["job"]["unramified_primes"] == [19]
    assert doc["dimension_total"] == sum(doc["dimensions_by_conductor"].values())
