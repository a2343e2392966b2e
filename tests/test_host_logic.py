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


@pytest.mark.parametrize("name,p", [("85085", 101), ("85085", 3), ("1062347", 29), ("11", 97)])      # primes not dividing the level
def test_isometry_third_row_identity(tables, port_oracle, name, p):
    """The algebra behind recover_row3 (csrc/tb_neighbor.cuh), checked on the oracle's neighbor records with
    exact integers: for a neighbor isometry S of representative Q with reduced form Q', S^T G S = p^2 G' and
    det S = +p^3, hence p G S = cof(S) G' and the third row of S follows from the other two,
        2c r3 = (r1 x r2) G' / p - g r1 - f r2,
    where (r1 x r2) G' is formed modulo 2^64 and divided by p through p^-1 mod 2^64 -- as the kernel does."""
    t = tables(name)
    rec = port_oracle(name).neighbor_records(p, False)          # [N][p+1][20]: line, form, isometry, r, spin
    mask = (1 << 64) - 1
    pinv = pow(p, -1, 1 << 64)

    def gram(q):
        a, b, c, f, g, h = (int(x) for x in q)
        return [[2 * a, h, g], [h, 2 * b, f], [g, f, 2 * c]]

    def signed(x):
        x &= mask
        return x - (1 << 64) if x >> 63 else x

    checked = 0
    for n in range(min(t.N, 12)):
        G = gram(t.q[n])
        g0, f0, c0 = int(t.q[n][4]), int(t.q[n][3]), int(t.q[n][2])
        for row in rec[n]:
            Gp = gram(row[3:9])
            S = [[int(row[9 + 3 * i + j]) for j in range(3)] for i in range(3)]
            # S^T G S = p^2 G'
            GS = [[sum(G[i][k] * S[k][j] for k in range(3)) for j in range(3)] for i in range(3)]
            StGS = [[sum(S[k][i] * GS[k][j] for k in range(3)) for j in range(3)] for i in range(3)]
            assert StGS == [[p * p * Gp[i][j] for j in range(3)] for i in range(3)]
            det = (S[0][0] * (S[1][1] * S[2][2] - S[1][2] * S[2][1]) - S[0][1] * (S[1][0] * S[2][2] - S[1][2] * S[2][0])
                   + S[0][2] * (S[1][0] * S[2][1] - S[1][1] * S[2][0]))
            assert det == p ** 3
            r1, r2 = S[0], S[1]
            x = [r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]]
            w = [sum(x[i] * Gp[i][j] for i in range(3)) & mask for j in range(3)]           # modulo 2^64
            tt = [signed(w[j] * pinv) - (g0 * r1[j] + f0 * r2[j]) for j in range(3)]
            assert all(v % (2 * c0) == 0 for v in tt)
            assert [v // (2 * c0) for v in tt] == S[2]
            checked += 1
    assert checked >= 24


@pytest.mark.parametrize("name,p", [("85085", 101), ("1062347", 29), ("11", 997)])
def test_fast_setup_closed_form(tables, port_oracle, name, p):
    """The closed form fast_setup (csrc/tb_neighbor.cuh) computes on the generic lane -- line with z = 1, g' != 0
    mod p: residues s1, s2 with the reference's truncate-then-shift rule, the neighbor's six coefficients with the
    two exact divisions, and the isometry [[u - s2, -s1 p, -p^2], [v, p, 0], [1, 0, 0]] -- restated with exact
    integers, reduced by the oracle's QuadForm::reduce and compared with the oracle's own neighbor records."""
    import ctypes
    t = tables(name)
    po = port_oracle(name)
    lib = po.lib
    lib.tbo_isotropic_vector.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    lib.tbo_reduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    rec = po.neighbor_records(p, False)
    base = po.base_points(p)
    pp, half = p * p, p >> 1
    generic = 0
    for n in range(min(t.N, 6)):
        q = np.ascontiguousarray(t.q[n], dtype=np.int64)
        b3 = np.ascontiguousarray(base[n], dtype=np.int64)
        a, b, c, f, g, h = (int(x) for x in q)
        vec = np.zeros(3, dtype=np.int64)
        for tt in range(p + 1):
            lib.tbo_isotropic_vector(q.ctypes.data, b3.ctypes.data, p, tt, vec.ctypes.data)
            vx, vy, vz = (int(x) for x in vec)
            assert [vx, vy, vz] == [int(x) for x in rec[n][tt][:3]]
            if vz != 1:
                continue
            u = vx - p if vx > half else vx
            v = vy - p if vy > half else vy
            au, hu, hv, bv = a * u, h * u, h * v, b * v
            temp = au + hv + g
            aa = c + u * temp + v * (bv + f)
            gg = -(temp + au)
            hh = f + hu + 2 * bv
            bb, cc, ff = b, a, -h
            if gg % p == 0:
                continue                                   # rotated neighbor: the integer path
            ginv = pow(gg % p, -1, p)
            c1 = (-hh * ginv) % p
            s1 = c1 - p if (hh > 0 and c1 >= p - half) else c1
            c2 = (((-aa) % pp) * ginv) % pp
            s2 = c2 - pp if c2 >= pp - (pp >> 1) else c2
            temp1 = cc * s1
            bb += s1 * (ff + temp1)
            ff += 2 * temp1
            hh1 = hh + s1 * gg
            temp2 = cc * s2
            num_a = aa + s2 * (gg + temp2)
            num_h = hh1 + s2 * ff
            assert num_a % pp == 0 and num_h % p == 0      # both divisions are exact (NeighborManager.h:338-341)
            form = np.array([num_a // pp, bb, cc * pp, ff * p, gg + 2 * temp2, num_h // p], dtype=np.int64)
            iso = np.array([u - s2, -s1 * p, -pp, v, p, 0, 1, 0, 0], dtype=np.int64)
            lib.tbo_reduce(form.ctypes.data, iso.ctypes.data, 0)
            assert np.array_equal(form, rec[n][tt][3:9]), (n, tt)
            assert np.array_equal(iso, rec[n][tt][9:18]), (n, tt)
            generic += 1
    assert generic >= 0.9 * min(t.N, 6) * (p + 1)
