"""Result persistence, format 1: the sqlite eigenvector database.

The reference keeps its results in one sqlite file (``src/sage/build_database.py``): a ``genus``
row per level (how the genus was built: ramified primes and seed, and how many rational
eigenvectors it has, -1 = not searched yet), the eigenvectors as comma-separated text, and one
``eigenvalues`` row per (vector, prime).  THE FILE is the contract -- the five SQL statements
below are the reference's (build_database.py:20-47), so either side opens the other's database --
and so are the names of the four calls a user makes (``get_genus``,
``get_genus_with_eigenvectors``, ``update_database_from_genus``, ``commit``).  What happens
behind them is written for this package: rows go through a handful of small query helpers, every
multi-row write is one sqlite transaction (an interrupted write rolls back by itself), and the
oldform pruning is a lookup by conductor instead of a double loop.
"""
from __future__ import annotations

import sqlite3

from .genus import BirchGenus

SCHEMA = (
    '''CREATE TABLE IF NOT EXISTS genus (
        level               INTEGER     PRIMARY KEY,
        ramified_primes     TEXT        NOT NULL,
        num_eigenvectors    INTEGER     NOT NULL DEFAULT -1,
        seed                INTEGER     NOT NULL
    )''',
    '''CREATE TABLE IF NOT EXISTS eigenvectors (
        id                  INTEGER     PRIMARY KEY,
        level               INTEGER     NOT NULL,
        conductor           INTEGER     NOT NULL,
        vector              TEXT        NOT NULL,
        FOREIGN KEY (level) REFERENCES genus (level)
    )''',
    '''CREATE TABLE IF NOT EXISTS eigenvalues (
        vector_id           INTEGER     NOT NULL,
        p                   INTEGER     NOT NULL,
        value               INTEGER     NOT NULL,
        UNIQUE (vector_id, p),
        FOREIGN KEY (vector_id) REFERENCES eigenvectors (id)
    )''',
    '''CREATE INDEX IF NOT EXISTS eigenvalues_index
        ON eigenvectors (level, conductor)''',
    '''CREATE INDEX IF NOT EXISTS eigenvectors_index
        ON eigenvalues (vector_id, p)''',
)

NOT_SEARCHED = -1                      # genus.num_eigenvectors before the eigenvector search has run
OLDFORM_EIGENVALUE_BOUND = 100         # eigenvalues computed before pruning, enough to tell the copies apart


def csv_of(values) -> str:
    """The text packing of the ``ramified_primes`` and ``vector`` columns (build_database.py:320-336)."""
    return ",".join(str(int(v)) for v in values)


def ints_of(text: str) -> list:
    return [int(tok) for tok in text.split(",")]


def oldform_copies(vectors, q: int) -> list:
    """Indices of the eigenvectors that are oldforms coming from level N / q.

    With an even number of prime divisors the largest one, q, stays unramified, and every newform of
    level N / q then shows up exactly twice in the search: once at a conductor d and once at d * q,
    with the same eigenvalues (build_database.py:112-152).  Both copies go.  Two vectors "agree" when
    their eigenvalues coincide at every prime both have."""
    by_conductor = {}
    for n, vec in enumerate(vectors):
        by_conductor.setdefault(int(vec["conductor"]), []).append(n)

    def agree(u, v):
        return all(v["aps"][p] == e for p, e in u["aps"].items() if p in v["aps"])

    drop = set()
    for n, vec in enumerate(vectors):
        cond = int(vec["conductor"])
        if cond % q:
            continue
        for m in by_conductor.get(cond // q, ()):
            if agree(vec, vectors[m]):
                drop.update((n, m))
    return sorted(drop)


class EigenvectorDatabase:
    """One sqlite file of genera, rational eigenvectors and eigenvalues; a context manager."""

    def __init__(self, filename, genus_factory=BirchGenus):
        self.filename = filename
        self.genus_factory = genus_factory              # tests inject a stub; the default builds the genus on the GPU
        self.conn = sqlite3.connect(filename)
        self.conn.row_factory = sqlite3.Row
        self.conn.execute("PRAGMA foreign_keys = 1")
        for statement in SCHEMA:
            self.conn.execute(statement)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def commit(self):
        self.conn.commit()

    def rollback(self):
        self.conn.rollback()

    def close(self):
        self.conn.close()

    # -- packing (static, as the reference exposes them) ---------------------------------------
    pack_vector = staticmethod(csv_of)
    unpack_vector = staticmethod(ints_of)
    unpack_ramified_primes = staticmethod(ints_of)

    @staticmethod
    def pack_ramified_primes(genus) -> str:
        return csv_of(genus.ramified_primes())

    @staticmethod
    def identify_duplicates(vectors, prime):
        return oldform_copies(vectors, int(prime))

    # -- row helpers ----------------------------------------------------------------------------
    def _genus_row(self, level: int):
        return self.conn.execute("SELECT * FROM genus WHERE level=?", (level,)).fetchone()

    def _insert_vectors(self, level: int, vectors) -> None:
        """Eigenvectors + the eigenvalues they carry + the genus row's count, as one transaction."""
        with self.conn:
            for vec in vectors:
                cur = self.conn.execute("INSERT INTO eigenvectors (level, conductor, vector) VALUES (?, ?, ?)",
                                        (level, int(vec["conductor"]), csv_of(vec["vector"])))
                vec["vector_id"] = cur.lastrowid
                self._insert_eigenvalues(vec, replace=False)
            self.conn.execute("UPDATE genus SET num_eigenvectors=? WHERE level=?", (len(vectors), level))

    def _insert_eigenvalues(self, vec, replace: bool) -> None:
        verb = "INSERT OR IGNORE" if replace else "INSERT"
        self.conn.executemany(f"{verb} INTO eigenvalues (vector_id, p, value) VALUES (?, ?, ?)",
                              [(vec["vector_id"], int(p), int(e)) for p, e in vec["aps"].items()])

    def _stored_vectors(self, level: int) -> list:
        rows = self.conn.execute("SELECT id, conductor, vector FROM eigenvectors WHERE level=? ORDER BY id", (level,)).fetchall()
        vectors = {row["id"]: dict(conductor=int(row["conductor"]), vector=ints_of(row["vector"]), aps={},
                                   vector_id=row["id"]) for row in rows}
        ids = list(vectors)
        for lo in range(0, len(ids), 500):              # sqlite caps the number of bound parameters
            chunk = ids[lo:lo + 500]
            marks = ", ".join("?" * len(chunk))
            for row in self.conn.execute(f"SELECT vector_id, p, value FROM eigenvalues WHERE vector_id IN ({marks})", chunk):
                vectors[row["vector_id"]]["aps"][int(row["p"])] = int(row["value"])
        return list(vectors.values())

    # -- the reference's calls ------------------------------------------------------------------
    def get_genus(self, level):
        """The genus of `level`: rebuilt from the stored ramified primes and seed when the level is known
        (so representatives come in the same order and stored vectors stay valid), else built fresh and
        registered (build_database.py:65-88)."""
        level = int(level)
        row = self._genus_row(level)
        if row is not None:
            return self.genus_factory(level, ramified_primes=ints_of(row["ramified_primes"]), seed=row["seed"])
        genus = self.genus_factory(level)
        self.conn.execute("INSERT INTO genus (level, ramified_primes, seed) VALUES (?, ?, ?)",
                          (level, self.pack_ramified_primes(genus), int(genus.seed())))
        return genus

    def get_genus_with_eigenvectors(self, level, precise=True, remove_oldforms=True):
        """The genus with its rational eigenvectors attached: searched for (and stored) on first contact,
        re-attached from the file afterwards (build_database.py:90-222)."""
        level = int(level)
        genus = self.get_genus(level)
        known = self._genus_row(level)["num_eigenvectors"]
        if known == NOT_SEARCHED:
            vectors = genus.rational_eigenvectors(precise=precise)
            genus.compute_eigenvalues_upto(OLDFORM_EIGENVALUE_BOUND, precise=precise)
            ramified_product = 1
            for p in genus.ramified_primes():
                ramified_product *= int(p)
            if remove_oldforms and ramified_product != level:
                doomed = oldform_copies(vectors, level // ramified_product)
                for n in reversed(doomed):
                    del vectors[n]
                if doomed:
                    genus.reset_eigenvector_manager()   # positions changed: the evaluation plan is rebuilt on next use
            self._insert_vectors(level, vectors)
        elif known > 0:
            for vec in self._stored_vectors(level):
                genus.add_eigenvector(vec)
        else:
            genus.eigenvectors = []                     # searched before, none found: nothing to look for again
        return genus

    def update_database_from_genus(self, genus, precise=True):
        """Store every eigenvalue the genus' eigenvectors have picked up since they were loaded
        (build_database.py:224-268).  The genus must be one this database handed out."""
        level = int(genus.level())
        row = self.conn.execute("SELECT num_eigenvectors FROM genus WHERE level=? AND seed=? AND ramified_primes=?",
                                (level, int(genus.seed()), self.pack_ramified_primes(genus))).fetchone()
        if row is None:
            raise LookupError(f"level {level} with this seed and these ramified primes is not in {self.filename}")
        if row["num_eigenvectors"] == NOT_SEARCHED:
            raise LookupError(f"the eigenvectors of level {level} have not been searched for yet")
        if row["num_eigenvectors"] == 0:
            return
        vectors = genus.rational_eigenvectors(precise=precise)
        missing = [n for n, vec in enumerate(vectors) if "vector_id" not in vec]
        if missing:
            raise LookupError(f"eigenvectors {missing} of level {level} carry no vector_id: they did not come from this database")
        with self.conn:
            for vec in vectors:
                self._insert_eigenvalues(vec, replace=True)
