"""Host-side twins of the reference's operator interface for the p-neighbor path.

``Genus``              <-> C++ ``Genus<R>``            (Genus.h:27-857)
``EigenvectorManager`` <-> C++ ``EigenvectorManager``  (Eigenvector.h:65-240)
``IsometrySequence``   <-> C++ ``IsometrySequence``    (IsometrySequence.h:16-102)
``BirchGenus``         <-> Cython ``BirchGenus``       (ternary_birch.pyx:124-841)

Same method names, argument meaning and error behaviour; every computation goes
through the C ABI of ``libtbirch.so`` (``cabi``) onto the GPU.  R = int64
(``precise=False``, wrap-around) or checked 128-bit (``precise=True``, replacing
GMP).  Genus *construction* is not part of the hot path: a ``Genus`` is made from
a ``GenusTable`` (see ``table.py``).
"""
from __future__ import annotations

import ctypes
from math import exp, lgamma, log

import numpy as np

from . import cabi
from .table import GenusTable, build_genus_table, load_genus_table, prime_symbols


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


class Eigenvector:
    """Eigenvector lifted to the full genus (Eigenvector.h:7-63)."""

    def __init__(self, data: np.ndarray, conductor_index: int):
        self._data = np.ascontiguousarray(data, dtype=np.int32)
        self._conductor_index = int(conductor_index)
        self._rep_index = -1

    def data(self) -> np.ndarray:
        return self._data

    def size(self) -> int:
        return int(self._data.size)

    def conductor_index(self) -> int:
        return self._conductor_index

    def rep_index(self, pos=None):
        if pos is None:
            return self._rep_index
        self._rep_index = int(pos)

    def __getitem__(self, pos):
        return int(self._data[pos])


class EigenvectorManager:
    """Eigenvector.h:65-240.  ``finalize`` picks, for every eigenvector, a genus
    representative at which its coordinate is nonzero (a set cover of the
    supports; the reference uses SetCover::METHOD_KINDA_GREEDY, any cover yields
    the same eigenvalues)."""

    def __init__(self):
        self.finalized = False
        self.dimension = 0
        self.eigenvectors: list[Eigenvector] = []
        self.indices: list[int] = []
        self.position_lut: list[list[int]] = []
        self.conductors: list[int] = []

    def add_eigenvector(self, vector: Eigenvector) -> None:
        if self.finalized:
            raise RuntimeError("Cannot add eigenvectors once finalized.")      # std::logic_error, Eigenvector.h:72
        if self.dimension > 0:
            if self.dimension != vector.size():
                raise ValueError("Eigenvector dimensions must match.")         # Eigenvector.h:79
        else:
            self.dimension = vector.size()
        self.eigenvectors.append(vector)

    def size(self) -> int:
        return len(self.eigenvectors)

    def __getitem__(self, index) -> Eigenvector:
        return self.eigenvectors[index]

    def finalize(self) -> None:
        if self.finalized:
            raise RuntimeError("Cannot finalize again.")                        # Eigenvector.h:99
        if not self.eigenvectors:
            self.finalized = True
            return
        support = np.stack([v.data() != 0 for v in self.eigenvectors])          # [V, N]
        if not support.any(axis=1).all():
            raise RuntimeError("Eigenvector is identically zero.")
        uncovered = np.ones(len(self.eigenvectors), dtype=bool)
        chosen = []
        while uncovered.any():                                                  # greedy set cover
            gain = support[uncovered].sum(axis=0)
            best = int(np.argmax(gain))
            chosen.append(best)
            uncovered &= ~support[:, best]
        self.indices = sorted(chosen)                                           # Eigenvector.h:160-161
        self.position_lut = [[] for _ in self.indices]
        for n, vec in enumerate(self.eigenvectors):                             # Eigenvector.h:170-185
            for pos, index in enumerate(self.indices):
                if vec.data()[index]:
                    vec.rep_index(index)
                    self.position_lut[pos].append(n)
                    break
        self.conductors = [v.conductor_index() for v in self.eigenvectors]
        self.finalized = True


class Genus:
    """Twin of ``Genus<R>``: owns the device copy of the genus table."""

    def __init__(self, table: GenusTable, precise: bool = False, stream: int | None = None):
        self.table = table
        self.precise = bool(precise)
        L = cabi.lib()
        self._keep = dict(
            primes=np.ascontiguousarray(table.primes, dtype=np.int64),
            conductors=np.ascontiguousarray(table.conductors, dtype=np.int64),
            dims=np.ascontiguousarray(table.dims, dtype=np.int64),
            q=np.ascontiguousarray(table.q, dtype=np.int64),
            s=np.ascontiguousarray(table.s, dtype=np.int64),
            sinv=np.ascontiguousarray(table.sinv, dtype=np.int64),
            scalar=np.ascontiguousarray(table.scalar, dtype=np.int64),
            lut=np.ascontiguousarray(table.lut, dtype=np.int32),
        )
        k = self._keep
        self._desc = cabi.GenusDesc(table.N, table.K, table.seed, table.disc, k["primes"].ctypes.data,
                                    k["conductors"].ctypes.data, k["dims"].ctypes.data, k["q"].ctypes.data,
                                    k["s"].ctypes.data, k["sinv"].ctypes.data, k["scalar"].ctypes.data,
                                    k["lut"].ctypes.data)
        self._h = ctypes.c_void_p()
        cabi.check(L.tb_genus_create(ctypes.byref(self._desc), ctypes.byref(self._h)))
        if stream is not None:
            self.set_stream(stream)

    # -- lifetime ---------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            cabi.lib().tb_genus_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int) -> None:
        cabi.check(cabi.lib().tb_genus_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def upload(self) -> None:
        """Re-send the genus table host -> device (bench e2e: inputs start on the host)."""
        cabi.check(cabi.lib().tb_genus_upload(self._h, ctypes.byref(self._desc)))

    def table_bytes(self) -> int:
        t = self.table
        return int((2 * 16 * 8 + 8) * t.N + 4 * t.lut.size)

    # -- Genus<R> surface --------------------------------------------------
    def size(self) -> int:                                   # Genus.h:306-309
        return self.table.N

    def seed(self) -> int:                                   # Genus.h:311-314
        return self.table.seed

    def dimension_map(self) -> dict:                         # Genus.h:316-325
        return {int(c): int(d) for c, d in zip(self.table.conductors, self.table.dims)}

    def _hecke(self, p: int, rep_begin=None, rep_end=None):
        L = cabi.lib()
        h = ctypes.c_void_p()
        if rep_begin is None:
            cabi.check(L.tb_hecke_compute(self._h, int(p), int(self.precise), ctypes.byref(h)))
        else:
            cabi.check(L.tb_hecke_compute_rows(self._h, int(p), int(self.precise), int(rep_begin),
                                               int(rep_end), ctypes.byref(h)))
        return h

    def hecke_matrix_sparse(self, p: int, rep_begin=None, rep_end=None) -> dict:
        """Genus<R>::hecke_matrix_sparse (Genus.h:341-348): {conductor: [data, indices, indptr]}."""
        L = cabi.lib()
        h = self._hecke(p, rep_begin, rep_end)
        try:
            # one read-back of all conductors into three arrays; the per-conductor results are views
            C = L.tb_hecke_num_conductors(h)
            nnz = [L.tb_hecke_nnz(h, k) for k in range(C)]
            rows = [L.tb_hecke_rows(h, k) for k in range(C)]
            data = np.empty(sum(nnz), dtype=np.int32)
            indices = np.empty(sum(nnz), dtype=np.int32)
            indptr = np.empty(sum(rows) + C, dtype=np.int32)
            cabi.check(L.tb_hecke_copy_sparse_all(h, _ptr(data), _ptr(indices), _ptr(indptr)))
            out, a, b = {}, 0, 0
            for k in range(C):
                out[int(L.tb_hecke_conductor(h, k))] = [data[a:a + nnz[k]], indices[a:a + nnz[k]],
                                                        indptr[b:b + rows[k] + 1]]
                a += nnz[k]
                b += rows[k] + 1
            return out
        finally:
            L.tb_hecke_destroy(h)

    def hecke_matrix_dense(self, p: int, rep_begin=None, rep_end=None) -> dict:
        """Genus<R>::hecke_matrix_dense (Genus.h:332-339): {conductor: row-major int32 rows x dim}."""
        L = cabi.lib()
        h = self._hecke(p, rep_begin, rep_end)
        try:
            out = {}
            for k in range(L.tb_hecke_num_conductors(h)):
                rows, dim = L.tb_hecke_rows(h, k), L.tb_hecke_dim(h, k)
                m = np.zeros((rows, dim), dtype=np.int32)
                cabi.check(L.tb_hecke_copy_dense(h, k, _ptr(m)))
                out[int(L.tb_hecke_conductor(h, k))] = m
            return out
        finally:
            L.tb_hecke_destroy(h)

    def hecke_device(self, p: int, rep_begin=None, rep_end=None) -> "HeckeResult":
        """T_p left resident on the device (bench `value`; multi-GPU gather)."""
        return HeckeResult(self._hecke(p, rep_begin, rep_end), owner=self)

    def eigenvector(self, vec, conductor: int) -> Eigenvector:
        """Genus<R>::eigenvector (Genus.h:350-390)."""
        t = self.table
        ks = np.nonzero(np.asarray(t.conductors) == int(conductor))[0]
        if ks.size == 0:
            raise ValueError("Invalid conductor.")
        k = int(ks[0])
        vec = np.asarray(vec, dtype=np.int32)
        if int(t.dims[k]) != vec.size:
            raise ValueError("Eigenvector has incorrect dimension.")
        lut = t.lut[k]
        full = np.zeros(t.N, dtype=np.int32)
        sel = lut != -1
        full[sel] = vec[lut[sel]]
        return Eigenvector(full, k)

    def eigenvalues(self, manager: EigenvectorManager, p: int) -> list:
        """Genus<R>::eigenvalues (Genus.h:392-421)."""
        V = manager.size()
        if V == 0:
            return []
        if not manager.finalized:
            raise RuntimeError("EigenvectorManager must be finalized.")
        vecs = np.ascontiguousarray(np.stack([v.data() for v in manager.eigenvectors]), dtype=np.int32)
        cond = np.asarray(manager.conductors, dtype=np.int64)
        rep = np.asarray([v.rep_index() for v in manager.eigenvectors], dtype=np.int64)
        out = np.zeros(V, dtype=np.int32)
        cabi.check(cabi.lib().tb_eigenvalues(self._h, int(p), int(self.precise), V, _ptr(vecs), _ptr(cond),
                                             _ptr(rep), _ptr(out)))
        return [int(x) for x in out]

    def neighbor_records(self, p: int, rep_begin: int = 0, rep_end=None) -> np.ndarray:
        """[rows, p+1, 20] int64: line[3], reduced form[6], isometry[9], r, spin (Genus.h:567-618)."""
        rep_end = self.table.N if rep_end is None else rep_end
        out = np.zeros((rep_end - rep_begin, int(p) + 1, 20), dtype=np.int64)
        cabi.check(cabi.lib().tb_neighbor_records(self._h, int(p), int(self.precise), int(rep_begin),
                                                  int(rep_end), _ptr(out)))
        return out

    def last_kernel_time(self):
        a, b, n = ctypes.c_float(), ctypes.c_float(), ctypes.c_int64()
        cabi.check(cabi.lib().tb_last_kernel_time(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
        return a.value, b.value, n.value

    INT_MODES = {-1: "none", 0: "int64-wrap", 1: "int64-bounded", 2: "int128-bounded", 3: "int128-checked"}

    def last_int_mode(self) -> str:
        """Integer mode the last neighbor kernel ran in (diagnostic)."""
        return self.INT_MODES[int(cabi.lib().tb_last_int_mode(self._h))]


class HeckeResult:
    """All 2^K matrices of one T_p, resident on the device until copied."""

    def __init__(self, handle, owner=None):
        self._h = handle
        self._owner = owner          # the result reads its genus (streams, copier): keep it alive

    def num_conductors(self) -> int:
        return int(cabi.lib().tb_hecke_num_conductors(self._h))

    def shape(self, k):
        L = cabi.lib()
        return int(L.tb_hecke_rows(self._h, k)), int(L.tb_hecke_dim(self._h, k)), int(L.tb_hecke_nnz(self._h, k))

    def conductor(self, k) -> int:
        return int(cabi.lib().tb_hecke_conductor(self._h, k))

    def row_begin(self, k) -> int:
        return int(cabi.lib().tb_hecke_row_begin(self._h, k))

    def wait_copies(self) -> None:
        """Block until every copy_sparse_to_async issued on this result has landed in its destination."""
        cabi.check(cabi.lib().tb_hecke_copy_wait(self._h))

    def transfer_bytes(self, k) -> int:
        """Bytes a host read-back of conductor k moves over PCIe (narrow 3-byte entries when they apply)."""
        return int(cabi.lib().tb_hecke_transfer_bytes(self._h, int(k)))

    def total_nnz(self) -> int:
        return sum(self.shape(k)[2] for k in range(self.num_conductors()))

    def copy_sparse_to(self, k, data_ptr: int, indices_ptr: int, indptr_ptr: int) -> None:
        cabi.check(cabi.lib().tb_hecke_copy_sparse(self._h, k, ctypes.c_void_p(data_ptr),
                                                   ctypes.c_void_p(indices_ptr), ctypes.c_void_p(indptr_ptr)))

    def copy_sparse_to_async(self, k, data_ptr: int, indices_ptr: int, indptr_ptr: int, cuda_stream: int) -> None:
        """Enqueue the copy on `cuda_stream` (device or pinned destinations) without synchronising."""
        cabi.check(cabi.lib().tb_hecke_copy_sparse_async(self._h, k, ctypes.c_void_p(data_ptr), ctypes.c_void_p(indices_ptr),
                                                         ctypes.c_void_p(indptr_ptr), ctypes.c_void_p(cuda_stream)))

    def packed_bytes(self):
        """(bytes, entry_bytes) of the one-buffer form the final gather ships (tb_hecke_pack)."""
        fmt = ctypes.c_int()
        n = int(cabi.lib().tb_hecke_packed_bytes(self._h, ctypes.byref(fmt)))
        return n, int(fmt.value)

    def pack_to(self, device_ptr: int, cuda_stream: int) -> None:
        """All conductors into one device buffer of packed_bytes()[0] bytes, enqueued on `cuda_stream`."""
        cabi.check(cabi.lib().tb_hecke_pack(self._h, ctypes.c_void_p(device_ptr), ctypes.c_void_p(cuda_stream)))

    def sparse(self, k):
        rows, _, nnz = self.shape(k)
        data = np.empty(nnz, dtype=np.int32)
        indices = np.empty(nnz, dtype=np.int32)
        indptr = np.empty(rows + 1, dtype=np.int32)
        self.copy_sparse_to(k, data.ctypes.data, indices.ctypes.data, indptr.ctypes.data)
        return data, indices, indptr

    def close(self) -> None:
        if self._h is not None and self._h.value:
            cabi.lib().tb_hecke_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IsometrySequence:
    """Twin of IsometrySequence<R,S,T> (IsometrySequence.h:16-102)."""

    def __init__(self, genus: Genus, p: int):
        self._genus = genus
        self._h = ctypes.c_void_p()
        cabi.check(cabi.lib().tb_isoseq_open(genus._h, int(p), int(genus.precise), ctypes.byref(self._h)))

    def done(self) -> bool:
        return bool(cabi.lib().tb_isoseq_done(self._h))

    def next(self) -> dict:
        iso = np.zeros(9, dtype=np.int64)
        den, src, dst = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        cabi.check(cabi.lib().tb_isoseq_next(self._h, _ptr(iso), ctypes.byref(den), ctypes.byref(src),
                                             ctypes.byref(dst)))
        return dict(isometry=iso.reshape(3, 3), denominator=den.value, src=src.value, dst=dst.value)

    def read(self, max_records: int) -> np.ndarray:
        rec = np.zeros((max_records, 12), dtype=np.int64)
        got = ctypes.c_int64()
        cabi.check(cabi.lib().tb_isoseq_read(self._h, int(max_records), _ptr(rec), ctypes.byref(got)))
        return rec[:got.value]

    def close(self) -> None:
        if self._h is not None and self._h.value:
            cabi.lib().tb_isoseq_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    if n % 2 == 0:
        return n == 2
    d = 3
    while d * d <= n:
        if n % d == 0:
            return False
        d += 2
    return True


class BirchGenus:
    """Twin of the Cython ``BirchGenus`` (ternary_birch.pyx:124-841) without Sage.

    The genus table replaces ``BirchGenus(level, ramified_primes, seed)``'s
    constructor work; everything downstream keeps the pyx's names, caching and
    error behaviour.  Sparse results are scipy ``csr_matrix`` objects when scipy
    is importable (as the pyx returns), else ``(data, indices, indptr)``.
    """

    def __init__(self, level, ramified_primes=None, seed=None):
        """``BirchGenus(level, ramified_primes=None, seed=None)`` as in the pyx (139-205): builds the
        quadratic form and enumerates the genus on the GPU.  A ``GenusTable`` (or a path to one)
        may be passed instead of the level to reuse an existing table."""
        if isinstance(level, GenusTable):
            table = level
        elif isinstance(level, (str, bytes)) or hasattr(level, "__fspath__"):
            table = load_genus_table(level)
        else:
            table = build_genus_table(int(level), ramified_primes, seed)
            self.ramified_primes_ = prime_symbols(int(level), ramified_primes)[1]
        if not hasattr(self, "ramified_primes_"):
            self.ramified_primes_ = None                      # not recorded in a bare table
        self.table = table
        self.level_ = table.level
        self.dims = {int(c): int(d) for c, d in zip(table.conductors, table.dims)}
        self.seed_ = table.seed
        self._genus = {}
        self.hecke = {}
        self.eigenvectors = None
        self._managers = {}

    def _get(self, precise: bool) -> Genus:
        if precise not in self._genus:
            self._genus[precise] = Genus(self.table, precise=precise)
        return self._genus[precise]

    def dimensions(self):
        return self.dims

    def dimension(self):
        return sum(self.dims.values())

    def level(self):
        return self.level_

    def seed(self):
        return self.seed_

    def ramified_primes(self):                                # pyx:236-237
        return self.ramified_primes_

    def next_good_prime(self, p):
        while True:
            p += 1
            while not _is_prime(p):
                p += 1
            if self.level_ % p != 0:
                return p

    def _estimated_density(self, p, conductor):              # pyx:788-816
        if p == 2:
            p = p + 1
        M = (p + 1) // 2
        N = self.dims[conductor]
        if N == 0:
            return 0.0
        if N == 1:
            return 1.0
        if conductor == 1:
            return 1.0 - exp(2 * M * (log(N - 1) - log(N)))
        lg2M1 = lgamma(2 * M + 1)
        log4 = log(4)
        logN1, logN = log(N - 1), log(N)
        density = 0.0
        for i in range(M):
            logprob = 2 * (M - i) * logN1 + lg2M1 - (2 * M * logN + 2 * lgamma(i + 1) + i * log4 +
                                                     lgamma(2 * M - 2 * i + 1))
            e = exp(logprob)
            if e == 0:
                break
            density += e
        return 1.0 - density

    def hecke_matrix(self, p, conductor, sparse=None, precise=True):   # pyx:571-622
        p = int(p)
        if not _is_prime(p):
            raise Exception("p is not prime.")
        if self.level_ % conductor != 0:
            raise Exception("Conductor must divide the level.")
        if self.level_ % p == 0:
            raise Exception("Cannot compute Hecke matrix at primes dividing the level.")
        if p in self.hecke:
            if conductor in self.hecke[p]:
                return self.hecke[p][conductor]
            raise Exception("No Hecke matrix associated to this conductor. How did this happen?")
        if sparse is None:
            sparse = self._estimated_density(p, conductor) <= 0.3
        g = self._get(bool(precise))
        self.hecke[p] = {}
        if sparse:
            try:
                from scipy.sparse import csr_matrix
            except ImportError:          # pragma: no cover
                csr_matrix = None
            for cond, (data, indices, indptr) in g.hecke_matrix_sparse(p).items():
                dim = self.dims[cond]
                self.hecke[p][cond] = (csr_matrix((data, indices, indptr), shape=(dim, dim))
                                       if csr_matrix else (data, indices, indptr))
        else:
            for cond, m in g.hecke_matrix_dense(p).items():
                self.hecke[p][cond] = m
        if conductor in self.hecke[p]:
            return self.hecke[p][conductor]
        raise Exception("No Hecke matrix associated to this conductor. How did this happen?")

    def rational_eigenvectors(self, precise=True, sparse=None, force=False, log=None):   # pyx:246-379
        """Rational eigenvectors of the Hecke algebra on every conductor's space, as dicts
        {'vector', 'aps', 'conductor'}.  Kernels are computed modulo primes on the GPU and lifted
        to Q with an exact check (ternary_birch_b200/eigen.py); Sage is not needed."""
        if not force and self.eigenvectors is not None:
            return self.eigenvectors
        from . import eigen
        self.eigenvectors, self.eigenvector_search_stats = eigen.rational_eigenvectors(
            self, precise=precise, sparse=sparse, log=log)
        self.reset_eigenvector_manager()
        return self.eigenvectors

    def reset_eigenvector_manager(self):                     # pyx:432-454
        self._managers = {}

    def add_eigenvector(self, eigenvector):                  # pyx:383-388
        if self.eigenvectors is None:
            self.eigenvectors = []
        e = dict(eigenvector)
        e.setdefault("aps", {})
        self.eigenvectors.append(e)
        self._managers = {}

    def _manager(self, precise: bool) -> EigenvectorManager:
        if precise not in self._managers:
            g = self._get(precise)
            m = EigenvectorManager()
            for entry in self.eigenvectors:
                m.add_eigenvector(g.eigenvector(entry["vector"], int(entry["conductor"])))
            m.finalize()
            self._managers[precise] = m
        return self._managers[precise]

    def compute_eigenvalues(self, p, precise=True):          # pyx:390-430
        p = int(p)
        if not _is_prime(p):
            raise Exception("p is not prime.")
        if self.level_ % p == 0:
            raise Exception("Cannot compute eigenvalues at primes dividing the level.")
        if self.eigenvectors is None:
            self.rational_eigenvectors()
        aps = self._get(bool(precise)).eigenvalues(self._manager(bool(precise)), p)
        for n, vec in enumerate(self.eigenvectors):
            vec["aps"][p] = aps[n]
        return list(aps)

    def compute_eigenvalues_upto(self, upper, precise=True, force=False):   # pyx:456-498
        ps = []
        p = 1
        while True:
            p = self.next_good_prime(p)
            if p <= upper:
                ps.append(p)
            else:
                break
        if self.eigenvectors is None:
            self.rational_eigenvectors()
        for p in ps:
            if not force and all(p in vec["aps"] for vec in self.eigenvectors):
                continue
            self.compute_eigenvalues(p, precise=precise)

    def isometry_sequence(self, p, precise=True):            # pyx:500-569
        p = int(p)
        if not _is_prime(p):
            raise Exception("p is not prime.")
        if self.level_ % p == 0:
            raise Exception("Cannot compute Hecke matrix at primes dividing the level.")
        seq = IsometrySequence(self._get(bool(precise)), p)
        try:
            while not seq.done():
                for rec in seq.read(4096):
                    yield dict(isometry=rec[:9].reshape(3, 3).copy(), denominator=int(rec[9]),
                               src=int(rec[10]), dst=int(rec[11]))
        finally:
            seq.close()

    def __repr__(self):                                      # pyx:818-822
        facs = " * ".join(str(int(p)) for p in self.table.primes)
        return """Birch genus with level {} = {}
Ramified primes = {}
Dimensions = {}
Seed = {}""".format(self.level_, facs, self.ramified_primes_, self.dims, self.seed_)

    def __reduce__(self):                                    # pyx:824-827
        return (BirchGenus, (self.level_, self.ramified_primes_, self.seed_), self.__getstate__())

    def __getstate__(self):                                  # pyx:829-835
        return dict(eigenvectors=list(self.eigenvectors) if self.eigenvectors is not None else None)

    def __setstate__(self, state):                           # pyx:837-841
        self.eigenvectors = list(state["eigenvectors"]) if state["eigenvectors"] is not None else None
