"""Result persistence, format 2: one JSON document per level.

Twin of the reference's batch driver ``scripts/generate_eigenvalues.py`` (Job / JobResult,
22-47; job generation, 97-120; the per-level computation, 132-166; ``<level>.json``,
223-225) without Sage.  The reference fans levels out over a process pool of CPU workers;
here one process drives one GPU and levels are dealt round-robin to ranks."""
from __future__ import annotations

import json
import operator
import time
from contextlib import contextmanager
from dataclasses import asdict, dataclass
from fractions import Fraction
from functools import reduce
from math import ceil

from .genus import BirchGenus


@dataclass
class Job:                                           # generate_eigenvalues.py:22-35
    level: int
    seed: int
    ramified_primes: list
    ramified_primes_product: int
    unramified_primes: list
    unramified_primes_product: int
    sturm_bound_of_ramified_primes_product: int
    eigenvalues_upto: int
    elapsed_seconds: float | None = None
    elapsed_seconds_genus: float | None = None
    elapsed_seconds_eigenvectors: float | None = None
    elapsed_seconds_eigenvalues: float | None = None


@dataclass
class JobResult:                                     # generate_eigenvalues.py:38-47
    job: Job
    dimensions_by_conductor: dict
    dimension_total: int
    rational_eigenvectors_count: int
    rational_eigenvectors: list

    def json(self, **kwargs) -> str:
        return json.dumps(asdict(self), **kwargs)


@dataclass
class Stopwatch:                                     # generate_eigenvalues.py:50-80
    elapsed_sum: int = 0
    start_mark: int | None = None

    def start(self):
        if self.start_mark is None:
            self.start_mark = time.monotonic_ns()

    def stop(self):
        if self.start_mark is not None:
            self.elapsed_sum += time.monotonic_ns() - self.start_mark
            self.start_mark = None

    def elapsed(self) -> float:
        total = self.elapsed_sum + (0 if self.start_mark is None else time.monotonic_ns() - self.start_mark)
        return total / 10**9

    @staticmethod
    @contextmanager
    def contextmanager():
        stopwatch = Stopwatch()
        stopwatch.start()
        try:
            yield stopwatch
        finally:
            stopwatch.stop()


def prime_divisors(n: int) -> list:
    out, d = [], 2
    while d * d <= n:
        if n % d == 0:
            out.append(d)
            while n % d == 0:
                n //= d
        d += 1
    if n > 1:
        out.append(n)
    return out


def is_squarefree(n: int, *, num_primes=None) -> bool:   # generate_eigenvalues.py:83-91
    primes = prime_divisors(n)
    if not primes or reduce(operator.mul, primes) != n:
        return False
    if num_primes is not None:
        return len(primes) == num_primes
    return True


def sturm_bound(level: int, weight: int = 2) -> int:
    """Sage's ``sturm_bound(level, weight)``: ceil(weight * [SL2(Z) : Gamma0(level)] / 12) with
    index level * prod (1 + 1/p).  (Stated from Sage's documentation; Sage is not in this image.)"""
    index = Fraction(level)
    for p in prime_divisors(level):
        index *= Fraction(p + 1, p)
    return int(ceil(index * weight / 12))


def generate_jobs(level: int, *, seed: int, upto=None) -> Job:   # generate_eigenvalues.py:94-120
    primes = sorted(prime_divisors(level))
    ramified = primes[:-1] if len(primes) % 2 == 0 else primes
    unramified = [primes[-1]] if len(primes) % 2 == 0 else []
    ramified_product = reduce(operator.mul, ramified)
    bound = sturm_bound(ramified_product, 2)
    upto = min(upto or bound, bound)
    return Job(level=level, seed=seed, ramified_primes=list(map(int, ramified)),
               ramified_primes_product=int(ramified_product), unramified_primes=list(map(int, unramified)),
               unramified_primes_product=int(reduce(operator.mul, unramified, 1)), eigenvalues_upto=int(upto),
               sturm_bound_of_ramified_primes_product=int(bound))


def normalize_vectors(vecs):                         # generate_eigenvalues.py:123-129
    for vec in vecs:
        vec["vector"] = tuple(int(coef) for coef in vec["vector"])
        vec["conductor"] = int(vec["conductor"])
        vec["aps"] = {int(p): int(e) for p, e in vec["aps"].items()}
    return vecs


def find_rational_eigenvectors(job: Job, genus_factory=BirchGenus) -> JobResult:   # generate_eigenvalues.py:132-166
    with Stopwatch.contextmanager() as stopwatch:
        with Stopwatch.contextmanager() as stopwatch_genus:
            genus = genus_factory(job.level, seed=job.seed)
        with Stopwatch.contextmanager() as stopwatch_eigenvectors:
            vecs = genus.rational_eigenvectors()
        with Stopwatch.contextmanager() as stopwatch_eigenvalues:
            genus.compute_eigenvalues_upto(job.eigenvalues_upto)
    job.elapsed_seconds = stopwatch.elapsed()
    job.elapsed_seconds_genus = stopwatch_genus.elapsed()
    job.elapsed_seconds_eigenvectors = stopwatch_eigenvectors.elapsed()
    job.elapsed_seconds_eigenvalues = stopwatch_eigenvalues.elapsed()
    return JobResult(job=job,
                     dimensions_by_conductor={int(c): int(d) for c, d in genus.dimensions().items()},
                     dimension_total=int(genus.dimension()), rational_eigenvectors_count=len(vecs),
                     rational_eigenvectors=normalize_vectors(vecs))


def load_result(path) -> dict:
    """Read a ``<level>.json`` written by either side; JSON turned the integer keys of
    ``aps`` / ``dimensions_by_conductor`` into strings, turn them back."""
    doc = json.load(open(path))
    doc["dimensions_by_conductor"] = {int(k): v for k, v in doc["dimensions_by_conductor"].items()}
    for vec in doc["rational_eigenvectors"]:
        vec["aps"] = {int(k): v for k, v in vec["aps"].items()}
    return doc
