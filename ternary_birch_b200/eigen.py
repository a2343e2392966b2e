"""Rational eigenvector search: the consumer of T_p.

Twin of ``BirchGenus.rational_eigenvectors`` (ternary_birch.pyx:246-379) without Sage.  The
reference finds candidate eigenvalues e of T_p inside the Hasse interval as roots of the
characteristic polynomial over GF(q), takes ``(T_p - e).right_kernel()`` over Q and splits
eigenspaces of dimension > 1 with T_p' at the next good prime, restricted to the eigenspace.

Here every kernel is computed modulo primes P < 2^29 on the GPU (``tb_modkernel_*``,
csrc/tb_linalg.cuh), lifted to Q by CRT + rational reconstruction and then VERIFIED exactly
over Z, so a result is a proof: the lifted basis consists of true kernel vectors, and the
kernel over Q cannot be larger than the kernel modulo P.  Testing every integer in the Hasse
interval is equivalent to the reference's charpoly-root filter (a rational eigenvalue is an
integer root of the characteristic polynomial, hence a root modulo q; a root that is no
eigenvalue has a trivial kernel and is dropped by the reference too).

What is pinned: each returned vector is the primitive integer eigenvector whose first
non-zero entry is positive -- what Sage's echelonised ZZ-kernel basis yields for a
one-dimensional eigenspace, independent of the path that isolated it -- with the same
``aps`` and ``conductor`` fields.  What is not (no Sage in this image): the ORDER of the
list, which in the reference follows Sage's ordering of GF(q) roots; this module uses
0, -1, ..., -hasse, hasse, ..., 1 (factors x + c sorted by c), documented as unpinned.
"""
from __future__ import annotations

import ctypes
from collections import deque
from math import floor, gcd, isqrt, sqrt

import numpy as np

from . import cabi

HASSE_MULTIPLIER = 50            # pyx:43 (only sizes the reference's GF(q); kept for reference)


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    d = 2
    while d * d <= n:
        if n % d == 0:
            return False
        d += 1
    return True


def lifting_primes():
    """Primes below 2^29, descending (the moduli of the GPU eliminations)."""
    n = (1 << 29) - 1
    while True:
        if _is_prime(n):
            yield n
        n -= 2


# ---------------------------------------------------------------------------
# GPU kernels modulo P
# ---------------------------------------------------------------------------
def _collect(handle):
    L = cabi.lib()
    try:
        dim, cols = int(L.tb_modkernel_dim(handle)), int(L.tb_modkernel_cols(handle))
        free = np.zeros(dim, dtype=np.int64)
        basis = np.zeros((dim, cols), dtype=np.uint32)
        if dim:
            cabi.check(L.tb_modkernel_free_cols(handle, free.ctypes.data_as(ctypes.c_void_p)))
            cabi.check(L.tb_modkernel_basis(handle, basis.ctypes.data_as(ctypes.c_void_p)))
        ms = float(L.tb_modkernel_ms(handle))
    finally:
        L.tb_modkernel_destroy(handle)
    return tuple(int(x) for x in free), basis, ms


def kernel_mod_csr(data, indices, indptr, n: int, shift: int, P: int):
    """Right kernel of (A - shift*I) mod P for an n x n CSR matrix -> (free_cols, basis, ms)."""
    data = np.ascontiguousarray(data, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    h = ctypes.c_void_p()
    cabi.check(cabi.lib().tb_modkernel_csr(data.ctypes.data_as(ctypes.c_void_p), indices.ctypes.data_as(ctypes.c_void_p),
                                           indptr.ctypes.data_as(ctypes.c_void_p), int(n), int(shift), int(P),
                                           ctypes.byref(h)))
    return _collect(h)


def kernel_mod_dense(M: np.ndarray, P: int):
    """Right kernel of a dense int64 matrix mod P -> (free_cols, basis, ms)."""
    M = np.ascontiguousarray(M, dtype=np.int64)
    rows, cols = M.shape
    h = ctypes.c_void_p()
    cabi.check(cabi.lib().tb_modkernel_dense(M.ctypes.data_as(ctypes.c_void_p), rows, cols, int(P), ctypes.byref(h)))
    return _collect(h)


# ---------------------------------------------------------------------------
# lifting to Q
# ---------------------------------------------------------------------------
def rational_reconstruct(a: int, m: int, bound: int):
    """a mod m -> (num, den) with |num| <= bound, 0 < den <= bound and num = a den (mod m), or None."""
    r0, r1, s0, s1 = m, a % m, 0, 1
    while r1 > bound:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        s0, s1 = s1, s0 - q * s1
    if s1 == 0 or abs(s1) > bound:
        return None
    if s1 < 0:
        r1, s1 = -r1, -s1
    if gcd(r1, s1) != 1:
        return None
    return r1, s1


def normalize_vector(v):
    """Primitive integer vector with a positive first non-zero entry (Sage's echelonised
    ZZ-basis of a rank-one module)."""
    g = 0
    for x in v:
        g = gcd(g, abs(int(x)))
    if g == 0:
        return [0 for _ in v]
    out = [int(x) // g for x in v]
    first = next(x for x in out if x)
    if first < 0:
        out = [-x for x in out]
    return out


def lift_vector(residues, m: int):
    """Integer vector proportional to the rational vector whose entries are `residues` mod m
    (common-denominator reconstruction), or None when m is still too small."""
    bound = isqrt(m // 2)
    half = m // 2
    D = 1
    row = np.array([int(x) for x in residues], dtype=object)
    nums = np.zeros(len(row), dtype=object)
    start = 0
    n = len(row)
    while start < n:
        y = (row[start:] * D) % m
        y = np.where(y > half, y - m, y)
        bad = np.nonzero(np.abs(y) > bound)[0]
        stop = int(bad[0]) if len(bad) else len(y)
        nums[start:start + stop] = y[:stop]
        start += stop
        if start >= n:
            break
        rr = rational_reconstruct(int((row[start] * D) % m), m, bound)
        if rr is None:
            return None
        num, den = rr
        nums[:start] = nums[:start] * den
        D *= den
        if D > bound:
            return None
        nums[start] = num
        start += 1
    return normalize_vector(list(nums))


def rational_kernel(kernel_mod_fn, verify_fn, max_primes: int = 40):
    """Q-basis (primitive integer vectors, one per free column of the reduced echelon form)
    of a kernel given modulo primes.

    kernel_mod_fn(P) -> (free_cols, basis uint32 [d][n], ms);  verify_fn(v) -> bool checks an
    integer vector exactly over Z.  Returns (vectors, stats)."""
    state = None
    used = []
    ms_total = 0.0
    for P in lifting_primes():
        if len(used) >= max_primes:
            break
        free, basis, ms = kernel_mod_fn(P)
        ms_total += ms
        used.append(P)
        if state is None or len(free) < len(state["free"]):
            state = dict(free=free, m=P, res=basis.astype(object))      # earlier primes were bad
        elif free != state["free"]:
            continue                                                    # this prime is bad
        else:
            m = state["m"]
            inv = pow(m % P, -1, P)
            delta = ((basis.astype(object) - state["res"]) * inv) % P
            state["res"] = state["res"] + delta * m
            state["m"] = m * P
        if len(state["free"]) == 0:
            return [], dict(primes=used, ms=ms_total)
        vecs = []
        for row in state["res"]:
            v = lift_vector(row, state["m"])
            if v is None or not verify_fn(v):
                vecs = None
                break
            vecs.append(v)
        if vecs is not None:
            return vecs, dict(primes=used, ms=ms_total)
    raise RuntimeError("rational_kernel: no verified lift after %d primes" % len(used))


# ---------------------------------------------------------------------------
# exact integer matrix-vector product
# ---------------------------------------------------------------------------
def csr_matvec_exact(A, v):
    """A v over Z for a scipy CSR integer matrix and a list of Python ints."""
    vmax = max((abs(int(x)) for x in v), default=0)
    absA = abs(A)
    rowmax = int(absA.sum(axis=1).max()) if A.shape[0] and A.nnz else 0
    if vmax * rowmax < (1 << 62):
        return [int(x) for x in A.astype(np.int64).dot(np.array(v, dtype=np.int64))]
    out = []
    for r in range(A.shape[0]):
        lo, hi = A.indptr[r], A.indptr[r + 1]
        out.append(sum(int(A.data[k]) * int(v[A.indices[k]]) for k in range(lo, hi)))
    return out


def is_eigenvector(A, v, e: int) -> bool:
    w = csr_matvec_exact(A, v)
    return all(int(a) == e * int(b) for a, b in zip(w, v))


# ---------------------------------------------------------------------------
# the search (pyx:246-379)
# ---------------------------------------------------------------------------
def candidate_eigenvalues(hasse: int):
    """Integers of the Hasse interval in the (unpinned, see module docstring) order the
    reference would enqueue its GF(q) roots."""
    return [0] + [-k for k in range(1, hasse + 1)] + list(range(hasse, 0, -1))


def _as_csr(A):
    from scipy.sparse import csr_matrix, issparse
    if issparse(A):
        return A.tocsr()
    return csr_matrix(np.asarray(A))


def rational_eigenvectors(genus, precise=True, sparse=None, log=None):
    """All rational eigenvectors of the Hecke algebra on every conductor's space: a list of
    dicts {'vector', 'aps', 'conductor'} as in pyx:337-343.  `genus` is a BirchGenus twin."""
    stats = dict(kernels=0, gpu_ms=0.0, primes=0)
    found = []
    p = genus.next_good_prime(1)
    hasse = int(floor(2 * sqrt(1.0 * p)))
    jobs = deque()
    for conductor in sorted(genus.dims):
        if genus.dims[conductor] == 0:
            continue
        A = _as_csr(genus.hecke_matrix(p, conductor, sparse=sparse, precise=precise))
        for e in candidate_eigenvalues(hasse):
            jobs.append(dict(p=p, e=e, aps={}, matrix=A, conductor=conductor, subspace=None))

    while jobs:
        job = jobs.popleft()
        p, e, A, conductor, S = job["p"], job["e"], job["matrix"], job["conductor"], job["subspace"]
        n = A.shape[0]
        if S is None:
            def kernel_mod(P, A=A, e=e, n=n):
                return kernel_mod_csr(A.data, A.indices, A.indptr, n, e, P)

            def verify(v, A=A, e=e):
                return is_eigenvector(A, v, e)
            vecs, st = rational_kernel(kernel_mod, verify)
            basis = vecs
        else:
            # kernel of (A - e) S^T in the coordinates of the subspace S (d x n integer rows)
            Sobj = np.array(S, dtype=object)
            A64 = A.astype(np.int64)

            def kernel_mod(P, Sobj=Sobj, A64=A64, e=e):
                Sp = (Sobj % P).astype(np.int64)                       # d x n residues
                M = (A64.dot(Sp.T) - (e % P) * Sp.T) % P               # n x d, exact in int64
                return kernel_mod_dense(M, P)

            def verify(k, Sobj=Sobj, A=A, e=e):
                w = list(np.array(k, dtype=object).dot(Sobj))
                return is_eigenvector(A, w, e)
            coords, st = rational_kernel(kernel_mod, verify)
            basis = [normalize_vector(list(np.array(k, dtype=object).dot(Sobj))) for k in coords]
        stats["kernels"] += 1
        stats["gpu_ms"] += st["ms"]
        stats["primes"] += len(st["primes"])
        dim = len(basis)
        if log:
            log("p=%d conductor=%d e=%d: kernel dimension %d (%d primes, %.1f ms on the GPU)"
                % (p, conductor, e, dim, len(st["primes"]), st["ms"]))
        if dim == 0:
            continue
        aps = dict(job["aps"])
        aps[p] = e
        if dim == 1:
            found.append(dict(vector=basis[0], aps=aps, conductor=conductor))
            continue
        # split with the Hecke operator at the next good prime (pyx:345-375)
        p2 = genus.next_good_prime(p)
        hasse2 = int(floor(2 * sqrt(1.0 * p2)))
        A2 = _as_csr(genus.hecke_matrix(p2, conductor, sparse=sparse, precise=precise))
        for e2 in candidate_eigenvalues(hasse2):
            jobs.append(dict(p=p2, e=e2, aps=aps, matrix=A2, conductor=conductor, subspace=basis))
    return found, stats
