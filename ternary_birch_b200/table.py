"""Genus table container and its on-disk format.

The table is the state the reference's ``Genus<R>`` holds after construction
(``Genus.h:423-434``; ``GenusRep``, ``Genus.h:12-26``).  Building it (genus
enumeration, ``Genus.h:37-246``) is outside this package's hot path: tables
come from a ``.npz`` file holding the flat int64 stream documented below, the
same layout ``tb_genus_desc`` (``include/tbirch.h``) takes.

stream layout (int64):
  magic "TBGENUS1", N, K, seed, disc,
  prime_divisors[K], conductors[2^K], dims[2^K],
  q[N][6], s[N][9], sinv[N][9], scalar[N], parent[N], p[N], num_auts[N],
  lut_positions[2^K][N]
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

GENUS_MAGIC = 0x3153554E45474254  # "TBGENUS1"


@dataclass
class GenusTable:
    N: int
    K: int
    seed: int
    disc: int
    primes: np.ndarray      # [K] int64      Genus::prime_divisors
    conductors: np.ndarray  # [2^K] int64    Genus::conductors
    dims: np.ndarray        # [2^K] int64    Genus::dims
    q: np.ndarray           # [N,6] int64    GenusRep::q
    s: np.ndarray           # [N,9] int64    GenusRep::s
    sinv: np.ndarray        # [N,9] int64    GenusRep::sinv
    scalar: np.ndarray      # [N] int64      my_pow(GenusRep::es)
    parent: np.ndarray      # [N] int64
    p: np.ndarray           # [N] int64
    num_auts: np.ndarray    # [N] int64      QuadForm::num_automorphisms(rep.q)
    lut: np.ndarray         # [2^K,N] int32  Genus::lut_positions

    @property
    def level(self) -> int:
        out = 1
        for p in self.primes:
            out *= int(p)
        return out

    def to_stream(self) -> np.ndarray:
        parts = [np.array([GENUS_MAGIC, self.N, self.K, self.seed, self.disc], dtype=np.int64),
                 self.primes, self.conductors, self.dims, self.q.ravel(), self.s.ravel(),
                 self.sinv.ravel(), self.scalar, self.parent, self.p, self.num_auts,
                 self.lut.astype(np.int64).ravel()]
        return np.concatenate([np.asarray(x, dtype=np.int64) for x in parts])

    @staticmethod
    def from_stream(stream) -> "GenusTable":
        a = np.asarray(stream, dtype=np.int64)
        if a.size < 5 or int(a[0]) != GENUS_MAGIC:
            raise ValueError("not a genus table stream")
        N, K, seed, disc = (int(x) for x in a[1:5])
        C = 1 << K
        pos = 5

        def take(n):
            nonlocal pos
            out = a[pos:pos + n].copy()
            pos += n
            return out

        primes, conductors, dims = take(K), take(C), take(C)
        q = take(6 * N).reshape(N, 6)
        s = take(9 * N).reshape(N, 9)
        sinv = take(9 * N).reshape(N, 9)
        scalar, parent, p, num_auts = take(N), take(N), take(N), take(N)
        lut = take(C * N).reshape(C, N).astype(np.int32)
        if pos != a.size:
            raise ValueError("genus table stream has trailing data")
        return GenusTable(N, K, seed, disc, primes, conductors, dims, q, s, sinv, scalar, parent,
                          p, num_auts, lut)


def load_genus_table(path) -> GenusTable:
    path = str(path)
    if path.endswith(".npz"):
        with np.load(path) as z:
            return GenusTable.from_stream(z["stream"])
    return GenusTable.from_stream(np.fromfile(path, dtype=np.int64))


def save_genus_table(table: GenusTable, path) -> None:
    path = str(path)
    if path.endswith(".npz"):
        np.savez_compressed(path, stream=table.to_stream())
    else:
        table.to_stream().tofile(path)


def _factor(n: int):
    out, d = [], 2
    while d * d <= n:
        e = 0
        while n % d == 0:
            n //= d
            e += 1
        if e:
            out.append((d, e))
        d += 1
    if n > 1:
        out.append((n, 1))
    return out


def prime_symbols(level: int, ramified_primes=None):
    """The (p, power, ramified) triples BirchGenus.__init__ builds (ternary_birch.pyx:146-167):
    without an explicit list, every prime is ramified when their number is odd, all but the
    largest otherwise."""
    facs = _factor(int(level))
    ps = [p for p, _ in facs]
    if ramified_primes is None:
        ram = ps[:-1] if len(ps) % 2 == 0 else ps[:]
    else:
        ram = [p for p in ramified_primes if p in ps]
    return [(p, e, p in ram) for p, e in facs], ram


def quad_form(symbols):
    """QuadForm<Z>::get_quad_form (QuadForm.cpp:444-777) through the C ABI."""
    import ctypes
    from . import cabi
    arr = (cabi.PrimeSymbol * len(symbols))(*[cabi.PrimeSymbol(int(p), int(e), int(bool(r))) for p, e, r in symbols])
    form = np.zeros(6, dtype=np.int64)
    cabi.check(cabi.lib().tb_quad_form(arr, len(symbols), ctypes.c_void_p(form.ctypes.data)))
    return [int(x) for x in form]


def build_genus_table(level: int, ramified_primes=None, seed=None, form=None) -> GenusTable:
    """Genus<Z>(q, symbols, seed) (Genus.h:37-246) through the C ABI: enumerate the genus of
    the given level on the GPU, in the reference's discovery order."""
    import ctypes
    from . import cabi
    symbols, _ = prime_symbols(level, ramified_primes)
    L = cabi.lib()
    arr = (cabi.PrimeSymbol * len(symbols))(*[cabi.PrimeSymbol(int(p), int(e), int(bool(r))) for p, e, r in symbols])
    q = np.asarray(form if form is not None else quad_form(symbols), dtype=np.int64)
    handle = ctypes.c_void_p()
    cabi.check(L.tb_genus_enumerate(ctypes.c_void_p(q.ctypes.data), arr, len(symbols), int(seed or 0), ctypes.byref(handle)))
    try:
        d = cabi.GenusDesc()
        parent, rep_p, auts = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        cabi.check(L.tb_table_desc(handle, ctypes.byref(d), ctypes.byref(parent), ctypes.byref(rep_p), ctypes.byref(auts)))
        N, K = int(d.N), int(d.K)
        C = 1 << K

        def arr64(ptr, n):
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(n,)).copy()

        lut = np.ctypeslib.as_array(ctypes.cast(d.lut, ctypes.POINTER(ctypes.c_int32)), shape=(C * N,)).copy()
        return GenusTable(N, K, int(d.seed), int(d.disc), arr64(d.primes, K), arr64(d.conductors, C), arr64(d.dims, C),
                          arr64(d.q, 6 * N).reshape(N, 6), arr64(d.s, 9 * N).reshape(N, 9),
                          arr64(d.sinv, 9 * N).reshape(N, 9), arr64(d.scalar, N), arr64(parent, N), arr64(rep_p, N),
                          arr64(auts, N), lut.reshape(C, N))
    finally:
        L.tb_table_destroy(handle)
