"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY.

Python access to the two checkers:

* ``RefBinary``  -- runs ``oracle/_ref/birch_ref`` (the UNMODIFIED reference
  compiled by ``oracle/Makefile``) and parses its int64 streams;
* ``PortOracle`` -- ctypes binding of ``oracle/_port/libtb_oracle.so`` (the plain-C
  restatement in ``oracle/port/``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Nothing in the product package
does.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import tempfile
from dataclasses import dataclass
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REF_BIN = ORACLE_DIR / "_ref" / "birch_ref"
PORT_LIB = ORACLE_DIR / "_port" / "libtb_oracle.so"

GENUS_MAGIC = 0x3153554E45474254  # "TBGENUS1"


def build(ref: bool = True, port: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    targets = []
    if port:
        targets.append("port")
    if ref and Path("/root/reference/src").is_dir():
        targets.append("ref")
    if targets:
        subprocess.run(["make", "-C", str(ORACLE_DIR), *targets], check=True,
                       stdout=subprocess.DEVNULL)


@dataclass
class GenusArrays:
    """The genus table as flat numpy arrays (layout of ref_driver.cpp:write_genus)."""
    N: int
    K: int
    seed: int
    disc: int
    primes: np.ndarray      # [K] int64
    conductors: np.ndarray  # [2^K] int64
    dims: np.ndarray        # [2^K] int64
    q: np.ndarray           # [N,6] int64
    s: np.ndarray           # [N,9] int64
    sinv: np.ndarray        # [N,9] int64
    scalar: np.ndarray      # [N] int64
    parent: np.ndarray      # [N] int64
    p: np.ndarray           # [N] int64
    num_auts: np.ndarray    # [N] int64
    lut: np.ndarray         # [2^K,N] int32


def parse_genus(stream: np.ndarray) -> GenusArrays:
    a = np.asarray(stream, dtype=np.int64)
    assert int(a[0]) == GENUS_MAGIC, "not a genus table"
    N, K, seed, disc = (int(x) for x in a[1:5])
    C = 1 << K
    pos = 5

    def take(n):
        nonlocal pos
        out = a[pos:pos + n].copy()
        pos += n
        return out

    primes = take(K)
    conductors = take(C)
    dims = take(C)
    q = take(6 * N).reshape(N, 6)
    s = take(9 * N).reshape(N, 9)
    sinv = take(9 * N).reshape(N, 9)
    scalar = take(N)
    parent = take(N)
    p = take(N)
    num_auts = take(N)
    lut = take(C * N).reshape(C, N).astype(np.int32)
    assert pos == a.size
    return GenusArrays(N, K, seed, disc, primes, conductors, dims, q, s, sinv, scalar,
                       parent, p, num_auts, lut)


def load_genus(path) -> GenusArrays:
    return parse_genus(np.fromfile(str(path), dtype=np.int64))


def parse_hecke(stream: np.ndarray, sparse: bool):
    """-> list over conductor index k of dicts (conductor, dim, data/indices/indptr | matrix)."""
    a = np.asarray(stream, dtype=np.int64)
    C = int(a[0])
    pos = 1
    out = []
    for _ in range(C):
        cond, dim = int(a[pos]), int(a[pos + 1])
        pos += 2
        if sparse:
            nnz = int(a[pos]); pos += 1
            data = a[pos:pos + nnz].astype(np.int32); pos += nnz
            indices = a[pos:pos + nnz].astype(np.int32); pos += nnz
            indptr = a[pos:pos + dim + 1].astype(np.int32); pos += dim + 1
            out.append(dict(conductor=cond, dim=dim, data=data, indices=indices, indptr=indptr))
        else:
            m = a[pos:pos + dim * dim].astype(np.int32).reshape(dim, dim); pos += dim * dim
            out.append(dict(conductor=cond, dim=dim, matrix=m))
    assert pos == a.size
    return out


class RefError(RuntimeError):
    def __init__(self, kind, what):
        super().__init__(f"{kind}: {what}")
        self.kind = kind
        self.what = what


class RefBinary:
    """Runs the compiled reference on one genus (level, ramified primes, seed)."""

    def __init__(self, level: int, seed: int = 1, ramified=None, binary: Path = REF_BIN):
        if not Path(binary).exists():
            raise FileNotFoundError(f"{binary} missing: run `make -C oracle ref`")
        self.binary = str(binary)
        self.level = int(level)
        self.seed = int(seed)
        self.ramified = ramified

    @staticmethod
    def available() -> bool:
        return REF_BIN.exists()

    def _run(self, commands, timeout=None):
        argv = [self.binary, "--level", str(self.level), "--seed", str(self.seed)]
        if self.ramified is not None:
            argv += ["--ramified", ",".join(str(p) for p in self.ramified)]
        argv += [str(c) for c in commands]
        proc = subprocess.run(argv, capture_output=True, text=True, timeout=timeout)
        lines = [json.loads(l) for l in proc.stdout.splitlines() if l.startswith("{")]
        if proc.returncode != 0:
            err = next((l for l in lines if "error" in l), None)
            if err:
                raise RefError(err["error"], err["what"])
            raise RuntimeError(f"birch_ref failed ({proc.returncode}): {proc.stderr}")
        return lines

    def _stream(self, commands_before_out):
        with tempfile.TemporaryDirectory() as d:
            out = os.path.join(d, "out.bin")
            self._run(list(commands_before_out) + [out], timeout=900)
            return np.fromfile(out, dtype=np.int64)

    def genus(self) -> GenusArrays:
        return parse_genus(self._stream(["genus"]))

    def write_genus(self, path) -> None:
        self._run(["genus", str(path)])

    def hecke(self, p: int, precise: bool, sparse: bool):
        mode = "z" if precise else "z64"
        fmt = "sparse" if sparse else "dense"
        return parse_hecke(self._stream(["hecke", p, mode, fmt]), sparse)

    def neighbor_records(self, p: int, precise: bool) -> np.ndarray:
        a = self._stream(["spin", p, "z" if precise else "z64"])
        N, pp = int(a[0]), int(a[1])
        return a[2:].reshape(N, pp + 1, 20)

    def isometry_sequence(self, p: int, precise: bool) -> np.ndarray:
        a = self._stream(["isoseq", p, "z" if precise else "z64"])
        return a[1:].reshape(int(a[0]), 12)

    def eigenvalues(self, vectors, primes, precise: bool) -> dict:
        """vectors: list of (conductor, coords); -> {p: [a_p per vector]}"""
        with tempfile.TemporaryDirectory() as d:
            vf = os.path.join(d, "vecs.bin")
            flat = [len(vectors)]
            for cond, coords in vectors:
                flat += [int(cond), len(coords), *[int(c) for c in coords]]
            np.asarray(flat, dtype=np.int64).tofile(vf)
            out = os.path.join(d, "out.bin")
            self._run(["eigen", vf, ",".join(str(p) for p in primes), "z" if precise else "z64", out])
            a = np.fromfile(out, dtype=np.int64)
        V, P = int(a[0]), int(a[1])
        body = a[2:].reshape(P, V + 1)
        return {int(row[0]): [int(x) for x in row[1:]] for row in body}

    def time_hecke(self, p: int, precise: bool, sparse: bool, repeat: int = 1, timeout=None):
        return self._run(["time", p, "z" if precise else "z64", "sparse" if sparse else "dense", repeat],
                         timeout=timeout)


class _CGenus(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int64), ("K", ctypes.c_int64), ("seed", ctypes.c_uint64),
                ("primes", ctypes.c_void_p), ("q", ctypes.c_void_p), ("s", ctypes.c_void_p),
                ("sinv", ctypes.c_void_p), ("scalar", ctypes.c_void_p), ("lut", ctypes.c_void_p),
                ("dims", ctypes.c_void_p)]


class _CCsr(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int64), ("nnz", ctypes.c_int64),
                ("data", ctypes.POINTER(ctypes.c_int32)), ("indices", ctypes.POINTER(ctypes.c_int32)),
                ("indptr", ctypes.POINTER(ctypes.c_int32))]


class PortError(RuntimeError):
    def __init__(self, code):
        names = {1: "overflow", 2: "not_found", 3: "bad_prime", 4: "nomem"}
        super().__init__(f"oracle port error {code} ({names.get(code, '?')})")
        self.code = code


class PortOracle:
    """ctypes binding of the plain-C restatement."""

    def __init__(self, genus: GenusArrays, lib: Path = PORT_LIB):
        if not Path(lib).exists():
            raise FileNotFoundError(f"{lib} missing: run `make -C oracle port`")
        self.lib = ctypes.CDLL(str(lib))
        self.lib.tbo_hash_form.restype = ctypes.c_uint64
        self.lib.tbo_last_error_position.restype = ctypes.c_int64
        self.g = genus
        # keep contiguous copies alive for the lifetime of the binding
        self._keep = dict(
            primes=np.ascontiguousarray(genus.primes, dtype=np.int64),
            q=np.ascontiguousarray(genus.q, dtype=np.int64),
            s=np.ascontiguousarray(genus.s, dtype=np.int64),
            sinv=np.ascontiguousarray(genus.sinv, dtype=np.int64),
            scalar=np.ascontiguousarray(genus.scalar, dtype=np.int64),
            lut=np.ascontiguousarray(genus.lut, dtype=np.int32),
            dims=np.ascontiguousarray(genus.dims, dtype=np.int64),
        )
        k = self._keep
        self.cg = _CGenus(genus.N, genus.K, genus.seed, k["primes"].ctypes.data, k["q"].ctypes.data,
                          k["s"].ctypes.data, k["sinv"].ctypes.data, k["scalar"].ctypes.data,
                          k["lut"].ctypes.data, k["dims"].ctypes.data)

    def _check(self, rc):
        if rc != 0:
            raise PortError(rc)

    def base_points(self, p: int) -> np.ndarray:
        out = np.zeros((self.g.N, 3), dtype=np.int64)
        self._check(self.lib.tbo_base_points(ctypes.byref(self.cg), ctypes.c_uint64(p),
                                             ctypes.c_void_p(out.ctypes.data)))
        return out

    def neighbor_records(self, p: int, precise: bool) -> np.ndarray:
        out = np.zeros((self.g.N, p + 1, 20), dtype=np.int64)
        self._check(self.lib.tbo_neighbor_records(ctypes.byref(self.cg), ctypes.c_uint64(p),
                                                  int(precise), ctypes.c_void_p(out.ctypes.data)))
        return out

    def hecke(self, p: int, precise: bool, sparse: bool = True):
        C = 1 << self.g.K
        arr = (_CCsr * C)()
        rc = self.lib.tbo_hecke_sparse(ctypes.byref(self.cg), ctypes.c_uint64(p), int(precise), arr)
        try:
            self._check(rc)
            out = []
            for k in range(C):
                m = arr[k]
                nnz, dim = int(m.nnz), int(m.dim)
                data = np.ctypeslib.as_array(m.data, shape=(nnz,)).copy() if nnz else np.zeros(0, np.int32)
                indices = np.ctypeslib.as_array(m.indices, shape=(nnz,)).copy() if nnz else np.zeros(0, np.int32)
                indptr = np.ctypeslib.as_array(m.indptr, shape=(dim + 1,)).copy()
                if sparse:
                    out.append(dict(conductor=int(self.g.conductors[k]), dim=dim, data=data,
                                    indices=indices, indptr=indptr))
                else:
                    dense = np.zeros((dim, dim), dtype=np.int32)
                    for r in range(dim):
                        lo, hi = indptr[r], indptr[r + 1]
                        dense[r, indices[lo:hi]] = data[lo:hi]
                    out.append(dict(conductor=int(self.g.conductors[k]), dim=dim, matrix=dense))
            return out
        finally:
            self.lib.tbo_csr_free(arr, ctypes.c_int64(C))

    def eigenvalues(self, p: int, precise: bool, vecs: np.ndarray, cond: np.ndarray, rep: np.ndarray):
        vecs = np.ascontiguousarray(vecs, dtype=np.int32)
        cond = np.ascontiguousarray(cond, dtype=np.int64)
        rep = np.ascontiguousarray(rep, dtype=np.int64)
        V = vecs.shape[0]
        out = np.zeros(V, dtype=np.int32)
        self._check(self.lib.tbo_eigenvalues(ctypes.byref(self.cg), ctypes.c_uint64(p), int(precise),
                                             ctypes.c_int64(V), ctypes.c_void_p(vecs.ctypes.data),
                                             ctypes.c_void_p(cond.ctypes.data), ctypes.c_void_p(rep.ctypes.data),
                                             ctypes.c_void_p(out.ctypes.data)))
        return out

    def isometry_sequence(self, p: int, precise: bool) -> np.ndarray:
        out = np.zeros((self.g.N * (p + 1), 12), dtype=np.int64)
        self._check(self.lib.tbo_isometry_sequence(ctypes.byref(self.cg), ctypes.c_uint64(p), int(precise),
                                                   ctypes.c_void_p(out.ctypes.data)))
        return out

    def reduce(self, q, s, precise: bool):
        q = np.ascontiguousarray(q, dtype=np.int64).copy()
        s = np.ascontiguousarray(s, dtype=np.int64).copy()
        self.lib.tbo_reduce(ctypes.c_void_p(q.ctypes.data), ctypes.c_void_p(s.ctypes.data), int(precise))
        return q, s

    def hash_form(self, q) -> int:
        q = np.ascontiguousarray(q, dtype=np.int64)
        return int(self.lib.tbo_hash_form(ctypes.c_void_p(q.ctypes.data)))
