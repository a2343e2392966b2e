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
