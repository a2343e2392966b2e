"""Result persistence, format 2: one JSON document per level.

The reference's batch driver (``scripts/generate_eigenvalues.py``) writes ``<level>.json`` for
every squarefree level of a range: which primes were taken as ramified, how far the
eigenvalues go (the Sturm bound of the ramified part), the dimensions of the conductor
spaces and the rational eigenvectors with their eigenvalues.  That document -- key names,
key order, value types -- is the contract (``DOCUMENT_KEYS`` / ``JOB_KEYS`` below, from
generate_eigenvalues.py:22-47); everything that produces it here is this package's own:
a level is planned (`plan_level`), run on the GPU through the ``BirchGenus`` twin
(`run_level`) with the phases timed by one `PhaseClock`, and written with `write_document`.
The reference fans levels out over a pool of CPU processes; here one process drives one
GPU and ``scripts/generate_eigenvalues.py`` deals the levels round-robin to ranks.
"""
from __future__ import annotations

import json
import time
from fractions import Fraction

from .genus import BirchGenus

# the reference's document, in the order its dataclasses serialise (the on-disk contract)
JOB_KEYS = ("level", "seed", "ramified_primes", "ramified_primes_product", "unramified_primes",
            "unramified_primes_product", "sturm_bound_of_ramified_primes_product", "eigenvalues_upto",
            "elapsed_seconds", "elapsed_seconds_genus", "elapsed_seconds_eigenvectors", "elapsed_seconds_eigenvalues")
DOCUMENT_KEYS = ("job", "dimensions_by_conductor", "dimension_total", "rational_eigenvectors_count",
                 "rational_eigenvectors")


def prime_divisors(n: int) -> list:
    """Distinct prime divisors of n >= 1, ascending (trial division: levels are below ~10^7)."""
    found, d = [], 2
    while d * d <= n:
        if n % d == 0:
            found.append(d)
            while n % d == 0:
                n //= d
        d += 1 if d == 2 else 2
    if n > 1:
        found.append(n)
    return found


def squarefree_primes(n: int):
    """The prime divisors of n when n > 1 is squarefree, else None."""
    primes = prime_divisors(n)
    product = 1
    for p in primes:
        product *= p
    return primes if primes and product == n else None


def is_squarefree(n: int, *, num_primes=None) -> bool:
    """Squarefree, and (optionally) with exactly `num_primes` prime divisors: the level filter of the batch run."""
    primes = squarefree_primes(n)
    return primes is not None and (num_primes is None or len(primes) == num_primes)


def sturm_bound(level: int, weight: int = 2) -> int:
    """ceil(weight * [SL2(Z) : Gamma0(level)] / 12), index = level * prod_{p | level} (1 + 1/p): what Sage's
    ``sturm_bound(level, weight)`` returns (stated from its documentation; Sage is not in this image)."""
    index = Fraction(level)
    for p in prime_divisors(level):
        index *= Fraction(p + 1, p)
    bound = index * weight / 12
    return -((-bound.numerator) // bound.denominator)


class LevelPlan:
    """What the batch run does at one squarefree level (generate_eigenvalues.py:94-120): with an even number
    of prime divisors the largest stays unramified, and eigenvalues are computed up to the Sturm bound of the
    ramified part (or `upto`, whichever is smaller)."""

    def __init__(self, level: int, seed: int, upto=None):
        primes = squarefree_primes(int(level))
        if primes is None:
            raise ValueError(f"level {level} is not squarefree")
        self.level, self.seed = int(level), int(seed)
        split = len(primes) - 1 if len(primes) % 2 == 0 else len(primes)
        self.ramified, self.unramified = primes[:split], primes[split:]
        self.ramified_product = 1
        for p in self.ramified:
            self.ramified_product *= p
        self.unramified_product = 1
        for p in self.unramified:
            self.unramified_product *= p
        self.sturm = sturm_bound(self.ramified_product, 2)
        self.upto = self.sturm if not upto else min(int(upto), self.sturm)

    def job_record(self, clock: "PhaseClock") -> dict:
        values = dict(level=self.level, seed=self.seed, ramified_primes=list(self.ramified),
                      ramified_primes_product=self.ramified_product, unramified_primes=list(self.unramified),
                      unramified_primes_product=self.unramified_product,
                      sturm_bound_of_ramified_primes_product=self.sturm, eigenvalues_upto=self.upto,
                      elapsed_seconds=clock.total(), elapsed_seconds_genus=clock.seconds("genus"),
                      elapsed_seconds_eigenvectors=clock.seconds("eigenvectors"),
                      elapsed_seconds_eigenvalues=clock.seconds("eigenvalues"))
        return {key: values[key] for key in JOB_KEYS}


def plan_level(level: int, *, seed: int, upto=None) -> LevelPlan:
    return LevelPlan(level, seed, upto)


class PhaseClock:
    """Wall-clock seconds per named phase: ``with clock.phase("genus"): ...``."""

    class _Phase:
        def __init__(self, clock, name):
            self.clock, self.name = clock, name

        def __enter__(self):
            self.t0 = time.perf_counter()

        def __exit__(self, *exc):
            self.clock.spent[self.name] = self.clock.spent.get(self.name, 0.0) + time.perf_counter() - self.t0
            return False

    def __init__(self):
        self.spent = {}

    def phase(self, name):
        return PhaseClock._Phase(self, name)

    def seconds(self, name) -> float:
        return float(self.spent.get(name, 0.0))

    def total(self) -> float:
        return float(sum(self.spent.values()))


def plain_vector(vec: dict) -> dict:
    """One rational eigenvector as JSON-ready Python values (generate_eigenvalues.py:123-129's shape)."""
    return dict(vector=[int(x) for x in vec["vector"]], aps={int(p): int(e) for p, e in vec["aps"].items()},
                conductor=int(vec["conductor"]))


def run_level(plan: LevelPlan, genus_factory=BirchGenus) -> dict:
    """Genus, rational eigenvectors, eigenvalues up to the plan's bound (generate_eigenvalues.py:132-166);
    returns the level's document."""
    clock = PhaseClock()
    with clock.phase("genus"):
        genus = genus_factory(plan.level, seed=plan.seed)
    with clock.phase("eigenvectors"):
        vectors = genus.rational_eigenvectors()
    with clock.phase("eigenvalues"):
        genus.compute_eigenvalues_upto(plan.upto)
    dims = {int(c): int(d) for c, d in genus.dimensions().items()}
    values = dict(job=plan.job_record(clock), dimensions_by_conductor=dims, dimension_total=int(genus.dimension()),
                  rational_eigenvectors_count=len(vectors), rational_eigenvectors=[plain_vector(v) for v in vectors])
    return {key: values[key] for key in DOCUMENT_KEYS}


def write_document(doc: dict, directory) -> str:
    """``<directory>/<level>.json`` (generate_eigenvalues.py:223-225)."""
    path = f"{directory}/{doc['job']['level']}.json"
    with open(path, "w") as f:
        json.dump(doc, f)
    return path


def load_result(path) -> dict:
    """Read a ``<level>.json`` written by either side; JSON turned the integer keys of ``aps`` and
    ``dimensions_by_conductor`` into strings, turn them back."""
    with open(path) as f:
        doc = json.load(f)
    doc["dimensions_by_conductor"] = {int(k): v for k, v in doc["dimensions_by_conductor"].items()}
    for vec in doc["rational_eigenvectors"]:
        vec["aps"] = {int(k): v for k, v in vec["aps"].items()}
    return doc
