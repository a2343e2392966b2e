"""CPU: host-side logic of the Python twins (no GPU, no library compute calls)."""
import numpy as np
import pytest

from ternary_birch_b200 import prime_symbols
from ternary_birch_b200.genus import BirchGenus, Eigenvector, EigenvectorManager, _is_prime
from ternary_birch_b200 import shard


def test_prime_symbols_default_ramification():
    """ternary_birch.pyx:151-156: all primes ramified for an odd count, all but the largest otherwise."""
    syms, ram = prime_symbols(11 * 13 * 17 * 19 * 23)
    assert ram == [11, 13, 17, 19, 23] and all(r for _, _, r in syms)
    syms, ram = prime_symbols(2 * 3 * 5 * 7)
    assert ram == [2, 3, 5] and syms[-1] == (7, 1, False)
    syms, ram = prime_symbols(5 ** 3 * 7, ramified_primes=[7, 11])
    assert syms == [(5, 3, False), (7, 1, True)] and ram == [7]


def test_eigenvector_manager_set_cover():
    m = EigenvectorManager()
    vecs = [[0, 1, 1, -1], [1, 0, 2, -2], [3, 0, 1, 0], [0, 0, 0, 5]]
    for v in vecs:
        m.add_eigenvector(Eigenvector(np.array(v, dtype=np.int32), 0))
    with pytest.raises(ValueError, match="Eigenvector dimensions must match."):
        m.add_eigenvector(Eigenvector(np.zeros(3, dtype=np.int32), 0))
    m.finalize()
    assert m.indices == sorted(m.indices)
    for v in m.eigenvectors:                      # every vector is evaluated where it is nonzero
        assert v.data()[v.rep_index()] != 0 and v.rep_index() in m.indices
    assert sorted(sum(m.position_lut, [])) == [0, 1, 2, 3]
    with pytest.raises(RuntimeError, match="Cannot finalize again."):
        m.finalize()
    with pytest.raises(RuntimeError, match="Cannot add eigenvectors once finalized."):
        m.add_eigenvector(Eigenvector(np.zeros(4, dtype=np.int32), 0))


def test_density_estimate_and_prime_helpers(tables):
    b = BirchGenus(tables("1062347"))              # from a table: no GPU needed until a compute call
    assert b.level() == 1062347 and b.dimensions()[1] == 2016
    assert b.next_good_prime(10) == 29 and b.next_good_prime(1) == 2
    d101 = b._estimated_density(101, 1)
    assert 0.04 < d101 < 0.06                      # ~102 neighbors spread over 2016 classes
    assert b._estimated_density(10007, 1) > 0.99   # -> dense path, as in the pyx heuristic
    assert b._estimated_density(101, 17 * 19) <= d101 + 1e-9
    assert _is_prime(65521) and not _is_prime(65535)
    with pytest.raises(Exception, match="p is not prime."):
        b.hecke_matrix(100, 1)
    with pytest.raises(Exception, match="Conductor must divide the level."):
        b.hecke_matrix(101, 7)
    with pytest.raises(Exception, match="Cannot compute Hecke matrix at primes dividing the level."):
        b.hecke_matrix(11, 1)


def test_prime_and_row_shards_cover_everything():
    primes = [p for p in range(2, 200) if _is_prime(p)]
    for world in (1, 2, 4, 8):
        dealt = shard.prime_shards(primes, world)
        assert sorted(sum(dealt, [])) == primes
        assert max(len(x) for x in dealt) - min(len(x) for x in dealt) <= 1
        blocks = shard.row_shards(10316, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == 10316
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
