"""GPU: the CUDA path, called through the C ABI, against the oracle port, the committed
golden vectors from the unmodified reference and (when oracle/_ref travelled) the
reference binary itself.  Integer work: every comparison is bit-exact."""
import json

import numpy as np
import pytest

from conftest import csr_digest, primes_upto

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def genus(tables):
    from ternary_birch_b200 import Genus
    cache = {}

    def get(name, precise):
        key = (name, precise)
        if key not in cache:
            cache[key] = Genus(tables(name), precise=precise)
        return cache[key]
    yield get
    for g in cache.values():
        g.close()


def assert_same_csr(got: dict, want: list, tag=""):
    assert sorted(got) == sorted(m["conductor"] for m in want)
    for m in want:
        data, indices, indptr = got[m["conductor"]]
        assert np.array_equal(indptr, m["indptr"]), (tag, m["conductor"], "indptr")
        assert np.array_equal(indices, m["indices"]), (tag, m["conductor"], "indices")
        assert np.array_equal(data, m["data"]), (tag, m["conductor"], "data")


def test_c1_every_entry(genus, hecke_small):
    """Config 1: BirchGenus(11), T_p for all p < 100, all conductors, every matrix entry."""
    for precise in (False, True):
        g = genus("11", precise)
        for p in primes_upto(100):
            if p == 11:
                continue
            got = g.hecke_matrix_sparse(p)
            for k, cond in enumerate((1, 11)):
                tag = f"L11_p{p}_{'z' if precise else 'z64'}_k{k}"
                data, indices, indptr = got[cond]
                assert np.array_equal(data, hecke_small[tag + "_data"]), tag
                assert np.array_equal(indices, hecke_small[tag + "_indices"]), tag
                assert np.array_equal(indptr, hecke_small[tag + "_indptr"]), tag


def test_even_level(genus, hecke_small):
    g = genus("210_s7", False)
    for p in (11, 13, 59):
        got = g.hecke_matrix_sparse(p)
        for k, cond in enumerate(g.table.conductors):
            tag = f"L210_p{p}_z64_k{k}"
            data, indices, indptr = got[int(cond)]
            assert np.array_equal(data, hecke_small[tag + "_data"]), tag
            assert np.array_equal(indices, hecke_small[tag + "_indices"]), tag
            assert np.array_equal(indptr, hecke_small[tag + "_indptr"]), tag


@pytest.mark.parametrize("name,p,precise", [
    ("85085", 2, False), ("85085", 2, True), ("85085", 3, False), ("85085", 19, True),
    ("85085", 101, False), ("85085", 101, True), ("85085", 997, False),
    ("1062347", 2, False), ("1062347", 3, True), ("1062347", 101, False), ("1062347", 101, True),
    ("1009091", 2, False), ("1009091", 107, False), ("1009091", 1009, False),
])
def test_hecke_vs_oracle_and_golden(genus, port_oracle, hecke_digests, name, p, precise):
    """Configs 2-4 (sampled primes): every CSR array equals the oracle's, and its digest equals
    the digest of the unmodified reference's output."""
    got = genus(name, precise).hecke_matrix_sparse(p)
    want = port_oracle(name).hecke(p, precise)
    assert_same_csr(got, want, (name, p, precise))
    key = f"L{name}_p{p}_{'z' if precise else 'z64'}"
    if key in hecke_digests:
        t = genus(name, precise).table
        dig = [csr_digest(c, d, *got[int(c)]) for c, d in zip(t.conductors, t.dims)]
        assert dig == hecke_digests[key]


def test_c2_named_call(genus, port_oracle):
    """Config 2: hecke_matrix(101, 17*19) plus all conductors, int64 and precise, sparse and dense."""
    want = {m["conductor"]: m for m in port_oracle("1062347").hecke(101, False, sparse=False)}
    for precise in (False, True):
        g = genus("1062347", precise)
        dense = g.hecke_matrix_dense(101)
        assert sorted(dense) == sorted(want)
        for cond, m in dense.items():
            assert m.shape == (want[cond]["dim"], want[cond]["dim"])
            assert np.array_equal(m, want[cond]["matrix"]), (precise, cond)
        assert dense[17 * 19].shape[0] == g.dimension_map()[17 * 19]
        # conductor-1 rows sum to p + 1
        assert (dense[1].sum(axis=1) == 102).all()


def test_neighbor_records(genus, records_small):
    """Per-neighbor parity: line, reduced form, isometry, class index, spinor value."""
    for key, want in records_small.items():
        level, p, mode, kind = key.split("_")
        if kind != "rec":
            continue
        name = {"L11": "11", "L210": "210_s7", "L85085": "85085"}[level]
        got = genus(name, mode == "z").neighbor_records(int(p[1:]))
        assert np.array_equal(got, want), key


def test_neighbor_records_vs_oracle_midsize(genus, port_oracle):
    for precise in (False, True):
        got = genus("1062347", precise).neighbor_records(101, 100, 164)
        want = port_oracle("1062347").neighbor_records(101, precise)[100:164]
        assert np.array_equal(got, want)


def test_isometry_sequence(genus, records_small):
    from ternary_birch_b200 import IsometrySequence
    for key, want in records_small.items():
        level, p, mode, kind = key.split("_")
        if kind != "iso":
            continue
        name = {"L11": "11", "L210": "210_s7", "L85085": "85085"}[level]
        seq = IsometrySequence(genus(name, mode == "z"), int(p[1:]))
        first = seq.next()
        assert first["isometry"].ravel().tolist() == want[0][:9].tolist()
        rest = seq.read(len(want))
        assert np.array_equal(rest, want[1:]), key
        assert seq.done()
        with pytest.raises(ValueError, match="No more isometries."):
            seq.next()
        seq.close()


def test_eigenvalues_c3(tables, golden_dir):
    """Config 3: rational eigenvectors of level 5*7*11*13*17, compute_eigenvalues_upto(1000)."""
    from ternary_birch_b200 import BirchGenus
    gold = json.load(open(golden_dir / "eigen_85085.json"))
    for precise, key in ((False, "aps_z64"), (True, "aps_z")):
        b = BirchGenus(tables("85085"))
        for v in gold["vectors"]:
            b.add_eigenvector(v)
        b.compute_eigenvalues_upto(1000, precise=precise)
        for p in gold["primes"]:
            got = [vec["aps"][p] for vec in b.eigenvectors]
            assert got == gold[key][str(p)], (precise, p)
    b = BirchGenus(tables("85085"))
    for v in gold["vectors"]:
        b.add_eigenvector(v)
    for p in (65521, 65537, 99991):
        assert b.compute_eigenvalues(p, precise=True) == gold["aps_z_big"][str(p)]


def test_eigenvalues_11a(tables, golden_dir):
    from ternary_birch_b200 import BirchGenus
    gold = json.load(open(golden_dir / "eigen_11.json"))
    b = BirchGenus(tables("11"))
    for v in gold["vectors"]:
        b.add_eigenvector(v)
    for p, want in gold["aps_z"].items():
        assert b.compute_eigenvalues(int(p), precise=True) == want, p


def test_errors(genus):
    g = genus("11", False)
    with pytest.raises(ValueError, match="Prime must not divide the discriminant."):
        g.hecke_matrix_sparse(11)
    with pytest.raises(ValueError, match="Invalid conductor."):
        g.eigenvector([1, 1], 7)
    with pytest.raises(ValueError, match="Eigenvector has incorrect dimension."):
        g.eigenvector([1, 1, 1], 1)


def test_row_shards_concatenate(genus):
    """Rows computed in shards (the multi-GPU partition) concatenate to the full matrix."""
    g = genus("85085", False)
    full = g.hecke_matrix_sparse(101)
    cuts = [0, 17, 64, 65, 130]
    parts = [g.hecke_matrix_sparse(101, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    for cond, (data, indices, indptr) in full.items():
        d = np.concatenate([pt[cond][0] for pt in parts])
        i = np.concatenate([pt[cond][1] for pt in parts])
        ptr = [np.zeros(1, dtype=np.int64)]
        off = 0
        for pt in parts:
            ptr.append(pt[cond][2][1:].astype(np.int64) + off)
            off += int(pt[cond][2][-1])
        assert np.array_equal(d, data) and np.array_equal(i, indices)
        assert np.array_equal(np.concatenate(ptr), indptr)


def test_int64_failure_modes(genus, golden_dir):
    """int64 mode at large p on the level-1062347 genus (tests/golden/errors.json holds what the
    unmodified reference raises): p = 20011 and p = 40009 overflow int64 in build_neighbor -> the
    reference's overflow_error text, the same exception here.  (The reference compiled WITHOUT the
    binding's <stdlib.h> prelude fails earlier at p = 20011 with "Key not found.": its abs(int64_t)
    then resolves to abs(int) and truncates f = 4399578151 in reduce's exit test, QuadForm.h:689;
    oracle/Makefile builds the checker the way the Cython binding is built)"""
    errors = json.load(open(golden_dir / "errors.json"))
    g = genus("1062347", False)
    for p in (40009, 20011):
        want = errors[f"L1062347_p{p}_z64"]
        assert want["kind"] == "overflow_error"
        with pytest.raises(OverflowError) as info:
            g.hecke_matrix_sparse(p)
        assert str(info.value) == want["what"]
    # the checked 128-bit mode computes the same prime without error
    gp = genus("1062347", True)
    data, indices, indptr = gp.hecke_matrix_sparse(20011)[1]
    assert (np.add.reduceat(data.astype(np.int64), indptr[:-1]) == 20012).all()


def test_large_prime_digest_and_modes_agree(genus, hecke_digests):
    """T_10007 on the level-1062347 genus: digest of the reference's int64 output, and the
    checked 128-bit mode gives the identical matrices (the reference's GMP run is bit-identical
    to its int64 run)."""
    key = "L1062347_p10007_z64"
    if key not in hecke_digests:
        pytest.skip("slow golden digests not generated")
    t = genus("1062347", False).table
    for precise in (False, True):
        got = genus("1062347", precise).hecke_matrix_sparse(10007)
        dig = [csr_digest(c, d, *got[int(c)]) for c, d in zip(t.conductors, t.dims)]
        assert dig == hecke_digests[key], precise


def test_size_independent_properties_full_size(genus):
    """Properties that hold at any size (C5: p = 65521, precise): conductor-1 rows sum to p + 1,
    columns are strictly ascending with no stored zeros, dense equals sparse on a row block,
    and T_p is self-adjoint for the automorphism weights: T[n][r] |Aut r| = T[r][n] |Aut n|
    (the relation the reference's dense path uses to mirror, Genus.h:820-842)."""
    g = genus("1062347", True)
    p = 65521
    csr = g.hecke_matrix_sparse(p)
    data, indices, indptr = csr[1]
    assert (np.add.reduceat(data.astype(np.int64), indptr[:-1]) == p + 1).all()
    assert (data != 0).all()
    for r in range(0, g.table.N, 97):
        row = indices[indptr[r]:indptr[r + 1]]
        assert (np.diff(row) > 0).all()
    block = g.hecke_matrix_dense(p, 500, 516)[1]
    for i, r in enumerate(range(500, 516)):
        want = np.zeros(g.table.N, dtype=np.int32)
        want[indices[indptr[r]:indptr[r + 1]]] = data[indptr[r]:indptr[r + 1]]
        assert np.array_equal(block[i], want)
    from scipy.sparse import csr_matrix
    N = g.table.N
    T = csr_matrix((data.astype(np.int64), indices, indptr), shape=(N, N))
    aut = g.table.num_auts.astype(np.int64)
    TA = T.multiply(aut[None, :]).tocsr()           # (T A)[n][r] = T[n][r] * aut[r]
    assert (TA != TA.T.tocsr()).nnz == 0            # T A is symmetric


def test_bench_workload_properties(genus, hecke_digests):
    """The bench workload itself (C4: N = 10316, T_9973, int64 and precise): every CSR array of all 8
    conductors hashes to the digest of the unmodified reference's output, conductor-1 rows sum
    to p + 1, T A is symmetric for A = diag(|Aut|), both integer modes agree entry for entry, and
    the row blocks of a 3-way representative shard concatenate to the full matrix."""
    from scipy.sparse import csr_matrix
    p = 9973
    g = genus("1009091", False)
    N = g.table.N
    full = g.hecke_matrix_sparse(p)
    # the unmodified reference's output for exactly this call (tests/golden/make_golden.py --bench)
    key = "L1009091_p9973_z64"
    if key in hecke_digests:
        t = g.table
        assert [csr_digest(c, d, *full[int(c)]) for c, d in zip(t.conductors, t.dims)] == hecke_digests[key]
    data, indices, indptr = full[1]
    assert (np.add.reduceat(data.astype(np.int64), indptr[:-1]) == p + 1).all()
    T = csr_matrix((data.astype(np.int64), indices, indptr), shape=(N, N))
    TA = T.multiply(g.table.num_auts.astype(np.int64)[None, :]).tocsr()
    assert (TA != TA.T.tocsr()).nnz == 0
    precise = genus("1009091", True).hecke_matrix_sparse(p)
    for cond in full:
        for a, b in zip(full[cond], precise[cond]):
            assert np.array_equal(a, b), cond
    from ternary_birch_b200 import shard
    parts = [g.hecke_matrix_sparse(p, a, b) for a, b in shard.row_shards(N, 3)]
    whole = shard.concat_row_shards(parts)
    for cond in (1, 97, 1009091):
        for a, b in zip(whole[cond], full[cond]):
            assert np.array_equal(a, b), cond


@pytest.mark.parametrize("name,p", [("85085", 101), ("1009091", 1009), ("1062347", 4099), ("11", 7)])
def test_precise_kernel_variants_agree(tables, port_oracle, monkeypatch, name, p):
    """precise = 1 runs in one of three integer modes chosen from magnitude bounds the host proves
    per prime: 64-bit (bounded), unchecked 128-bit (bounded), overflow-checked 128-bit.  All three
    must give the reference's records, isometries and matrices bit for bit."""
    from ternary_birch_b200 import Genus, IsometrySequence
    t = tables(name)
    rows = min(t.N, 24)
    small = t.N * (p + 1) <= 200000
    got = {}
    for env, want_mode in ((None, "int64-bounded"), ("TBIRCH_FORCE_WIDE", "int128-bounded"),
                           ("TBIRCH_FORCE_CHECKED", "int128-checked")):
        monkeypatch.delenv("TBIRCH_FORCE_WIDE", raising=False)
        monkeypatch.delenv("TBIRCH_FORCE_CHECKED", raising=False)
        if env:
            monkeypatch.setenv(env, "1")
        g = Genus(t, precise=True)
        try:
            csr = g.hecke_matrix_sparse(p)
            mode = g.last_int_mode()
            rec = g.neighbor_records(p, 0, rows)
            iso = None
            if small:
                seq = IsometrySequence(g, p)
                iso = seq.read(t.N * (p + 1))
                seq.close()
        finally:
            g.close()
        assert mode == want_mode, (env, mode)
        got[env] = (csr, rec, iso)
    for env, (csr, rec, iso) in got.items():
        assert np.array_equal(rec, got[None][1]), env
        for cond in csr:
            for a, b in zip(csr[cond], got[None][0][cond]):
                assert np.array_equal(a, b), (env, cond)
        if small:
            assert np.array_equal(iso, got[None][2]), env
    oracle = port_oracle(name)
    assert_same_csr(got[None][0], oracle.hecke(p, True), name)
    if small:
        assert np.array_equal(got[None][1], oracle.neighbor_records(p, True)[:rows])
        assert np.array_equal(got[None][2], oracle.isometry_sequence(p, True))


@pytest.mark.parametrize("name,p,want_narrow", [("85085", 997, True), ("85085", 65521, False), ("1062347", 1009, True),
                                                 ("1009091", 107, True)])
def test_narrow_readback_matches_plain(tables, monkeypatch, name, p, want_narrow):
    """Host read-backs of large matrices travel as uint16 columns + int8 entries and are widened by
    host threads; the arrays must equal the plain int32 copy, into pageable and into pinned memory,
    and the plain path must be taken when an entry does not fit 8 bits (level 85085, p = 65521:
    ~500 neighbors per class)."""
    import torch
    from ternary_birch_b200 import Genus
    t = tables(name)
    got = {}
    for narrow in (False, True):
        monkeypatch.delenv("TBIRCH_NO_NARROW_COPY", raising=False)
        monkeypatch.delenv("TBIRCH_FORCE_NARROW_COPY", raising=False)
        monkeypatch.setenv("TBIRCH_NARROW_MIN_NNZ", "1")
        # pin the format: left alone, the library picks it per rank from measured delivery rates
        monkeypatch.setenv("TBIRCH_FORCE_NARROW_COPY" if narrow else "TBIRCH_NO_NARROW_COPY", "1")
        g = Genus(t, precise=True)
        try:
            res = g.hecke_device(p)
            rows, dim, nnz = res.shape(0)
            used = res.transfer_bytes(0) < 8 * nnz + 4 * (rows + 1)
            assert used == (narrow and want_narrow)
            got[narrow] = {k: res.sparse(k) for k in range(res.num_conductors())}
            if narrow:
                # asynchronous variant into pinned memory on a side stream
                side = torch.cuda.Stream()
                for k in (0, res.num_conductors() - 1):
                    rows, dim, nnz = res.shape(k)
                    d = torch.full((max(nnz, 1),), -7, dtype=torch.int32).pin_memory()
                    i = torch.full((max(nnz, 1),), -7, dtype=torch.int32).pin_memory()
                    ptr = torch.zeros(rows + 1, dtype=torch.int32).pin_memory()
                    torch.cuda.synchronize()
                    res.copy_sparse_to_async(k, d.data_ptr(), i.data_ptr(), ptr.data_ptr(), side.cuda_stream)
                    side.synchronize()
                    res.wait_copies()
                    assert np.array_equal(d.numpy()[:nnz], got[True][k][0])
                    assert np.array_equal(i.numpy()[:nnz], got[True][k][1])
                    assert np.array_equal(ptr.numpy(), got[True][k][2])
            res.close()
        finally:
            g.close()
    for k in got[False]:
        for a, b in zip(got[False][k], got[True][k]):
            assert np.array_equal(a, b), k


@pytest.mark.parametrize("env", [dict(TBIRCH_TWO_PASS_ROWS="1"), dict(TBIRCH_TWO_PASS_ROWS="1", TBIRCH_NARROW_MIN_NNZ="1"),
                                 dict(TBIRCH_ROWS_SCRATCH="1"), dict(TBIRCH_ROWS_SCRATCH="1", TBIRCH_TWO_PASS_ROWS="1"),
                                 dict(TBIRCH_NARROW_MIN_NNZ="1", TBIRCH_NARROW_MAX_RATIO="100000"),
                                 dict(TBIRCH_TWO_PASS_ROWS="1", TBIRCH_NARROW_MIN_NNZ="1", TBIRCH_NARROW_MAX_RATIO="100000"),
                                 dict(TBIRCH_ROWS_ACC="1"), dict(TBIRCH_ROWS_SORT="1")])
def test_row_kernel_variants_agree(tables, port_oracle, monkeypatch, env):
    """The row phase has a fused single-pass form and a count/scan/fill form, each with int32 or narrow
    output.  With the narrow attempt forced at p = 16381 >> N = 130 an entry exceeds 8 bits, so the fused form
    must redo the row pass in int32 and the two-pass form must pick int32 after its count pass.  All variants
    equal the oracle.  TBIRCH_ROWS_SCRATCH forces the form the kernel takes when a row's sort state exceeds
    one SM's shared memory (more than ~20000 classes at p > ~30000): the state in a global-memory slot per
    CTA, CTAs walking the rows persistently.  The single-pass form is the byte-counter kernel (rows_acc_kernel)
    on dense rows and the sorting kernel (rows_kernel) on sparse ones; TBIRCH_ROWS_ACC / TBIRCH_ROWS_SORT take
    each of them through all three primes."""
    from ternary_birch_b200 import Genus
    for k in ("TBIRCH_TWO_PASS_ROWS", "TBIRCH_NARROW_MIN_NNZ", "TBIRCH_NARROW_MAX_RATIO", "TBIRCH_NO_NARROW_COPY",
              "TBIRCH_ROWS_SCRATCH", "TBIRCH_ROWS_ACC", "TBIRCH_ROWS_SORT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.delenv("TBIRCH_FORCE_NARROW_COPY", raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    if "TBIRCH_NARROW_MIN_NNZ" in env:
        monkeypatch.setenv("TBIRCH_FORCE_NARROW_COPY", "1")   # the narrow attempt itself is under test, not the policy
    t = tables("85085")
    g = Genus(t, precise=True)
    try:
        for p in (19, 997, 16381):
            res = g.hecke_device(p)
            got = {int(res.conductor(k)): res.sparse(k) for k in range(res.num_conductors())}
            dense0 = g.hecke_matrix_dense(p)[1]
            res.close()
            assert_same_csr(got, port_oracle("85085").hecke(p, True), (env, p))
            data, indices, indptr = got[1]
            want = np.zeros((t.N, t.N), dtype=np.int32)
            for r in range(t.N):
                want[r, indices[indptr[r]:indptr[r + 1]]] = data[indptr[r]:indptr[r + 1]]
            assert np.array_equal(dense0, want), (env, p)
            if p == 16381:
                assert np.abs(data).max() > 127          # the case the int32 fallback exists for
    finally:
        g.close()


@pytest.mark.parametrize("two_pass", [False, True])
def test_row_kernel_global_scratch_walks_rows(tables, port_oracle, monkeypatch, two_pass):
    """The global-memory form of the row kernel with more rows (2016) than resident CTAs (296), so every
    CTA walks several rows and the look-back chains across walks."""
    from ternary_birch_b200 import Genus
    monkeypatch.setenv("TBIRCH_ROWS_SCRATCH", "1")
    if two_pass:
        monkeypatch.setenv("TBIRCH_TWO_PASS_ROWS", "1")
    else:
        monkeypatch.delenv("TBIRCH_TWO_PASS_ROWS", raising=False)
    g = Genus(tables("1062347"), precise=False)
    try:
        for p in (3, 101):
            assert_same_csr(g.hecke_matrix_sparse(p), port_oracle("1062347").hecke(p, False), (two_pass, p))
    finally:
        g.close()


@pytest.mark.parametrize("name,p", [("85085", 101), ("1062347", 1009), ("1009091", 107), ("11", 65521), ("85085", 9973), ("11", 7)])
@pytest.mark.parametrize("env", [None, "TBIRCH_NO_F32_TAIL", "TBIRCH_NO_ISO32", "TBIRCH_NO_FAST_SETUP", "TBIRCH_NO_PROVEN_BOUNDS"])
def test_reduce_loop_forms(tables, port_oracle, monkeypatch, name, p, env):
    """Setup and reduce loop have several exact forms chosen per launch from host-proven bounds: the
    neighbor set up in binary64 on the generic lane (fast_setup) or in integers on every lane
    (TBIRCH_NO_FAST_SETUP); the loop's form part in binary64 with a binary32 tail or binary64 to the end
    (TBIRCH_NO_F32_TAIL); its isometry in int32 ring arithmetic or in binary64 with lazy signs
    (TBIRCH_NO_ISO32); per-lane magnitude tests instead of proven bounds (TBIRCH_NO_PROVEN_BOUNDS).
    Each must reproduce the oracle's reduced forms AND isometries (neighbor records), matrices and
    isometry stream, in both integer modes."""
    from ternary_birch_b200 import Genus, IsometrySequence
    for v in ("TBIRCH_NO_F32_TAIL", "TBIRCH_NO_ISO32", "TBIRCH_NO_FAST_SETUP", "TBIRCH_NO_PROVEN_BOUNDS"):
        monkeypatch.delenv(v, raising=False)
    if env:
        monkeypatch.setenv(env, "1")
    t = tables(name)
    oracle = port_oracle(name)
    rows = min(t.N, 32)
    for precise in (False, True):
        g = Genus(t, precise=precise)
        try:
            assert np.array_equal(g.neighbor_records(p, 0, rows), oracle.neighbor_records(p, precise)[:rows]), precise
            assert_same_csr(g.hecke_matrix_sparse(p), oracle.hecke(p, precise), (name, p, precise, env))
            if t.N * (p + 1) <= 200000:
                seq = IsometrySequence(g, p)
                assert np.array_equal(seq.read(t.N * (p + 1)), oracle.isometry_sequence(p, precise))
                seq.close()
        finally:
            g.close()


@pytest.mark.parametrize("name,p", [("85085", 101), ("1062347", 101), ("85085", 997)])
def test_integer_reduce_fallback(tables, port_oracle, monkeypatch, name, p):
    """The reduction normally runs in the exact FP64 loop; the literal integer loop (eisenstein_reduce,
    QuadForm.h:530-691 statement by statement) is what out-of-range launches fall back to.
    TBIRCH_NO_FP64_REDUCE sends every neighbor through it: records (reduced form AND isometry),
    Hecke matrices and the isometry stream must still equal the oracle's, in both integer modes."""
    from ternary_birch_b200 import Genus, IsometrySequence
    monkeypatch.setenv("TBIRCH_NO_FP64_REDUCE", "1")
    t = tables(name)
    oracle = port_oracle(name)
    rows = min(t.N, 48)
    for precise in (False, True):
        g = Genus(t, precise=precise)
        try:
            assert np.array_equal(g.neighbor_records(p, 0, rows), oracle.neighbor_records(p, precise)[:rows]), precise
            assert_same_csr(g.hecke_matrix_sparse(p), oracle.hecke(p, precise), (name, p, precise))
            if t.N * (p + 1) <= 200000:
                seq = IsometrySequence(g, p)
                assert np.array_equal(seq.read(t.N * (p + 1)), oracle.isometry_sequence(p, precise))
                seq.close()
        finally:
            g.close()


def test_checked_128_bit_mode_raises_the_reference_overflow(tables):
    """precise=True replaces GMP by overflow-checked 128-bit arithmetic.  At p = 2^31 - 1 the second
    shear of build_neighbor forms a * s2^2 with |s2| up to p^2 = 2^62 (NeighborManager.h:325-329), past
    127 bits for any representative with a >= 8: the checked mode must notice and raise the reference's
    overflow_error text (NeighborManager.h:350-352) instead of returning wrapped values.  (GMP itself
    would succeed: documented deviation)  Reached through eigenvalues(), the only
    call that makes p + 1 = 2^31 neighbors of ONE representative."""
    from ternary_birch_b200 import EigenvectorManager, Genus
    t = tables("1062347")
    n = int(np.argmax(t.q[:, 0]))
    assert t.q[n, 0] >= 8 and t.lut[0][n] != -1
    g = Genus(t, precise=True)
    try:
        vec = np.zeros(int(t.dims[0]), dtype=np.int32)
        vec[int(t.lut[0][n])] = 1
        man = EigenvectorManager()
        man.add_eigenvector(g.eigenvector(vec, 1))
        man.finalize()
        with pytest.raises(OverflowError) as info:
            g.eigenvalues(man, 2147483647)
        assert str(info.value) == "An overflow has occurred. The p-neighbor's discriminant does not match the original."
        assert g.last_int_mode() == "int128-checked"
        # the same call at a prime the 128-bit pipeline does hold still works afterwards
        assert len(g.eigenvalues(man, 65537)) == 1
    finally:
        g.close()


def test_failing_calls_release_device_memory(tables):
    """A call that ends in the reference's overflow / "Key not found." error must give back the device
    buffers it reserved (GBs at large p): repeated failures leave the free-memory figure flat."""
    import torch
    from ternary_birch_b200 import Genus
    g = Genus(tables("1062347"), precise=False)
    try:
        free = []
        for _ in range(6):
            with pytest.raises((OverflowError, ValueError)):
                g.hecke_matrix_sparse(40009)
            torch.cuda.synchronize()
            free.append(torch.cuda.mem_get_info()[0])
        # the stream-ordered pool keeps what it once allocated; it must not keep growing
        assert free[-1] >= free[1] - (64 << 20), free
    finally:
        g.close()

