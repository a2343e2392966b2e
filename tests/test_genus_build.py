"""GPU: genus construction against the genus tables written by the
unmodified reference (tests/golden/genus_*.npz): the form from get_quad_form, the order of
the representatives, every isometry to / from the mother form, scalars, parents, |Aut|,
dimensions and lut_positions, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("11", 11, 1), ("210_s7", 210, 7), ("85085", 85085, 1), ("1062347", 1062347, 1), ("1009091", 1009091, 1)]


@pytest.mark.parametrize("name,level,seed", CASES)
def test_enumeration_matches_reference_table(tables, name, level, seed):
    from ternary_birch_b200 import build_genus_table, prime_symbols, quad_form
    want = tables(name)
    symbols, _ = prime_symbols(level)
    assert quad_form(symbols) == want.q[0].tolist()
    got = build_genus_table(level, seed=seed)
    assert (got.N, got.K, got.seed, got.disc) == (want.N, want.K, want.seed, want.disc)
    for field in ("primes", "conductors", "dims", "q", "parent", "p", "num_auts", "scalar", "s", "sinv", "lut"):
        assert np.array_equal(getattr(got, field), getattr(want, field)), field


def test_known_forms():
    from ternary_birch_b200 import prime_symbols, quad_form
    assert quad_form(prime_symbols(11 * 13 * 17 * 19 * 23)[0]) == [1, 187, 1467, -187, 0, 0]
    assert quad_form(prime_symbols(5 * 7 * 11 * 13 * 17)[0]) == [1, 130, 172, -65, -1, 0]
    assert quad_form(prime_symbols(97 * 101 * 103)[0]) == [1, 495, 515, -103, 0, 0]
    with pytest.raises(ValueError, match="Must specify an odd number of ramified primes."):
        quad_form([(11, 1, False)])


def test_birch_genus_from_level_end_to_end(hecke_small):
    """BirchGenus(11) as in the reference's README: constructor, T_p, eigenvalues."""
    from ternary_birch_b200 import BirchGenus
    g = BirchGenus(11, seed=1)
    assert g.dimensions() == {1: 2, 11: 0}
    T = g.hecke_matrix(3, 1, precise=False, sparse=False)
    assert T.tolist() == [[2, 2], [3, 1]]
    g.add_eigenvector(dict(vector=[2, -3], conductor=1))
    assert g.compute_eigenvalues(7) == [-2]


@pytest.mark.parametrize("level,seed", [(7 * 13 * 31, 3), (2 * 3 * 5 * 7 * 11, 5), (3 * 7 * 11 * 13, 4), (5 * 5 * 5, 4)])
def test_enumeration_vs_reference_binary(level, seed):
    """Further levels (even, non-squarefree), when the compiled reference travelled."""
    from oracle.pyoracle import RefBinary
    if not RefBinary.available():
        pytest.skip("oracle/_ref not present")
    from ternary_birch_b200 import build_genus_table
    want = RefBinary(level, seed).genus()
    got = build_genus_table(level, seed=seed)
    assert got.N == want.N
    for field in ("primes", "conductors", "dims", "q", "parent", "p", "num_auts", "scalar", "s", "sinv", "lut"):
        assert np.array_equal(getattr(got, field), getattr(want, field)), field


def test_randomised_levels_end_to_end_vs_reference_binary():
    """A deterministic sweep over further levels (odd / even, 1-4 prime divisors, varying seeds):
    build the genus here, then compare T_p (both integer modes), an isometry sequence and the genus
    table itself with the unmodified reference run on the same box."""
    from oracle.pyoracle import RefBinary
    if not RefBinary.available():
        pytest.skip("oracle/_ref not present")
    from ternary_birch_b200 import Genus, IsometrySequence, build_genus_table
    rng = np.random.default_rng(20261019)
    pool = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47]
    done = 0
    while done < 6:
        k = int(rng.integers(1, 5))
        ps = sorted(int(x) for x in rng.choice(pool, size=k, replace=False))
        level = int(np.prod(ps))
        if level > 40000:
            continue
        seed = int(rng.integers(1, 1000))
        ref = RefBinary(level, seed)
        want = ref.genus()
        got = build_genus_table(level, seed=seed)
        for field in ("q", "s", "sinv", "scalar", "lut", "dims", "num_auts"):
            assert np.array_equal(getattr(got, field), getattr(want, field)), (level, seed, field)
        good = [p for p in (2, 3, 5, 7, 11, 13) if level % p][:3]
        for precise in (False, True):
            g = Genus(got, precise=precise)
            for p in good:
                mine = g.hecke_matrix_sparse(p)
                for m in ref.hecke(p, precise, True):
                    d, i, ptr = mine[m["conductor"]]
                    assert np.array_equal(d, m["data"]) and np.array_equal(i, m["indices"]) and np.array_equal(ptr, m["indptr"]), (level, seed, p, precise)
            seq = IsometrySequence(g, good[0])
            rec = seq.read(got.N * (good[0] + 1))
            assert np.array_equal(rec, ref.isometry_sequence(good[0], precise)), (level, seed, precise)
            seq.close()
            g.close()
        done += 1


@pytest.mark.parametrize("level", [2, 3, 5, 7, 13, 37])
def test_class_number_one_and_two(level):
    """Smallest genera (one or two classes; level 2 makes p = 2 a bad prime)."""
    from ternary_birch_b200 import BirchGenus
    g = BirchGenus(level, seed=1)
    n = g.dimensions()[1]
    assert n == (2 if level == 37 else 1)
    for p in (2, 3, 5, 101):
        if level % p == 0:
            with pytest.raises(Exception):
                g.hecke_matrix(p, 1, precise=False)
            continue
        for precise in (False, True):
            g.hecke = {}
            T = g.hecke_matrix(p, 1, sparse=False, precise=precise)
            assert T.shape == (n, n) and (T.sum(axis=1) == p + 1).all()
            g.hecke = {}
            S = g.hecke_matrix(p, 1, sparse=True, precise=precise)
            assert (np.asarray(S.todense()) == T).all()
    recs = list(g.isometry_sequence(3 if level != 3 else 5))
    assert len(recs) == n * ((3 if level != 3 else 5) + 1)
