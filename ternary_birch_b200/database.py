"""Result persistence, format 1: the sqlite eigenvector database.

Twin of ``EigenvectorDatabase`` (reference: src/sage/build_database.py:15-336) without Sage:
the same three tables, indices, column names and text packing, so a database written by
either side is readable by the other.  Sage ``Integer`` / ``vector`` values become Python
``int`` / ``list``.  The genus behind it is the GPU-backed ``BirchGenus`` twin."""
from __future__ import annotations

import operator
import sqlite3
from functools import reduce

from .genus import BirchGenus


class EigenvectorDatabase:
    # build_database.py:20-47 -- verbatim schema (it is the on-disk contract)
    CREATE_TABLE_GENUS = '''CREATE TABLE IF NOT EXISTS genus (
        level               INTEGER     PRIMARY KEY,
        ramified_primes     TEXT        NOT NULL,
        num_eigenvectors    INTEGER     NOT NULL DEFAULT -1,
        seed                INTEGER     NOT NULL
    )'''

    CREATE_TABLE_EIGENVECTORS = '''CREATE TABLE IF NOT EXISTS eigenvectors (
        id                  INTEGER     PRIMARY KEY,
        level               INTEGER     NOT NULL,
        conductor           INTEGER     NOT NULL,
        vector              TEXT        NOT NULL,
        FOREIGN KEY (level) REFERENCES genus (level)
    )'''

    CREATE_TABLE_EIGENVALUES = '''CREATE TABLE IF NOT EXISTS eigenvalues (
        vector_id           INTEGER     NOT NULL,
        p                   INTEGER     NOT NULL,
        value               INTEGER     NOT NULL,
        UNIQUE (vector_id, p),
        FOREIGN KEY (vector_id) REFERENCES eigenvectors (id)
    )'''

    CREATE_INDEX_EIGENVECTORS = '''CREATE INDEX IF NOT EXISTS eigenvalues_index
        ON eigenvectors (level, conductor)'''

    CREATE_INDEX_EIGENVALUES = '''CREATE INDEX IF NOT EXISTS eigenvectors_index
        ON eigenvalues (vector_id, p)'''

    def __init__(self, filename, genus_factory=BirchGenus):
        self.filename = filename
        self.genus_factory = genus_factory          # tests inject cached genus tables here
        self.conn = self.create_connection(self.filename)
        self.conn.row_factory = sqlite3.Row
        self.create_database(self.conn)

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        self.conn.close()

    # build_database.py:65-88
    def get_genus(self, level):
        cur = self.conn.cursor()
        cur.execute("SELECT * FROM genus WHERE level=?", (int(level),))
        result = cur.fetchone()
        if result is None:
            genus = self.genus_factory(level)
            values = (int(level), self.pack_ramified_primes(genus), int(genus.seed()))
            cur.execute("INSERT INTO genus (level, ramified_primes, seed) VALUES (?, ?, ?)", values)
        else:
            assert level == result['level']
            genus = self.genus_factory(level, ramified_primes=self.unpack_ramified_primes(result['ramified_primes']),
                                       seed=result['seed'])
        cur.close()
        return genus

    # build_database.py:90-222
    def get_genus_with_eigenvectors(self, level, precise=True, remove_oldforms=True):
        genus = self.get_genus(level)
        level = int(level)
        cur = self.conn.cursor()
        cur.execute("SELECT num_eigenvectors FROM genus WHERE level=?", (level,))
        num_eigenvectors = cur.fetchone()['num_eigenvectors']
        if num_eigenvectors < 0:
            evecs = genus.rational_eigenvectors(precise=precise)
            genus.compute_eigenvalues_upto(100, precise=precise)
            if remove_oldforms:
                # an even number of prime factors leaves the largest unramified: oldforms then
                # occur exactly twice (build_database.py:112-152)
                old_level = reduce(operator.mul, genus.ramified_primes())
                if old_level != level:
                    q = level // old_level
                    removal = sorted(self.identify_duplicates(evecs, q))
                    for n in removal[::-1]:
                        del evecs[n]
                    if len(removal) > 0:
                        genus.reset_eigenvector_manager()
            try:
                for vec in evecs:
                    values = (level, int(vec['conductor']), self.pack_vector(vec['vector']))
                    cur.execute("INSERT INTO eigenvectors (level, conductor, vector) VALUES (?, ?, ?)", values)
                    vector_id = cur.lastrowid
                    vec['vector_id'] = vector_id
                    eigenvalues = [(vector_id, int(p), int(e)) for p, e in vec['aps'].items()]
                    cur.executemany("INSERT INTO eigenvalues (vector_id, p, value) VALUES (?, ?, ?)", eigenvalues)
                cur.execute("UPDATE genus SET num_eigenvectors=? WHERE level=?", (len(evecs), level))
            except KeyboardInterrupt:
                self.conn.rollback()
                raise KeyboardInterrupt("Rolling back database to previous commit.")
        elif num_eigenvectors > 0:
            cur.execute("SELECT * FROM eigenvectors WHERE level=?", (level,))
            results = cur.fetchall()
            assert len(results) > 0
            vector_ids = []
            vectors = dict()
            for result in results:
                vector_id = result['id']
                vector_ids.append(vector_id)
                vectors[vector_id] = dict(conductor=int(result['conductor']),
                                          vector=self.unpack_vector(result['vector']), aps=dict(),
                                          vector_id=vector_id)
            cur.execute("SELECT * FROM eigenvalues WHERE vector_id IN ({0})".format(", ".join("?" for _ in vector_ids)),
                        vector_ids)
            for result in cur.fetchall():
                vectors[result['vector_id']]['aps'][int(result['p'])] = int(result['value'])
            for vec in vectors.items():
                genus.add_eigenvector(vec[1])
        else:
            genus.eigenvectors = []                 # known to have none: no search needed
        cur.close()
        return genus

    # build_database.py:224-268
    def update_database_from_genus(self, genus, precise=True):
        level = int(genus.level())
        seed = int(genus.seed())
        ramified_primes = self.pack_ramified_primes(genus)
        cur = self.conn.cursor()
        cur.execute("SELECT * FROM genus WHERE level=? AND seed=? AND ramified_primes=?", (level, seed, ramified_primes))
        result = cur.fetchone()
        if result is None:
            raise Exception("No genus found in the database associated to this level, seed, and set of ramified primes.")
        if result['num_eigenvectors'] < 0:
            raise Exception("Unknown number of eigenvectors associated to this genus.")
        if result['num_eigenvectors'] == 0:
            return
        eigenvectors = genus.rational_eigenvectors(precise=precise)
        try:
            for vec in eigenvectors:
                if 'vector_id' not in vec:
                    raise Exception("No vector id associated to eigenvector within genus... TODO: Resolve this "
                                    "automatically instead of throwing an exception.")
                vector_id = vec['vector_id']
                eigenvalues = [(vector_id, int(p), int(value)) for p, value in vec['aps'].items()]
                cur.executemany("INSERT OR IGNORE INTO eigenvalues (vector_id, p, value) VALUES (?, ?, ?)", eigenvalues)
        except KeyboardInterrupt:
            self.conn.rollback()
            raise KeyboardInterrupt("Rolling back database to previous commit.")
        cur.close()

    def commit(self):
        self.conn.commit()

    def rollback(self):
        self.conn.rollback()

    def close(self):
        self.conn.close()

    @staticmethod
    def create_connection(filename):
        return sqlite3.connect(filename)

    @staticmethod
    def create_database(conn):
        c = conn.cursor()
        c.execute("PRAGMA foreign_keys = 1")
        c.execute(EigenvectorDatabase.CREATE_TABLE_GENUS)
        c.execute(EigenvectorDatabase.CREATE_TABLE_EIGENVECTORS)
        c.execute(EigenvectorDatabase.CREATE_TABLE_EIGENVALUES)
        c.execute(EigenvectorDatabase.CREATE_INDEX_EIGENVECTORS)
        c.execute(EigenvectorDatabase.CREATE_INDEX_EIGENVALUES)

    # build_database.py:300-315
    @staticmethod
    def identify_duplicates(vectors, prime):
        removal = []
        for n, vec1 in enumerate(vectors):
            cond = vec1['conductor']
            if cond % prime != 0:
                continue
            for m, vec2 in enumerate(vectors):
                if vec2['conductor'] * prime != cond:
                    continue
                match = True
                for p, value in vec1['aps'].items():
                    if match and p in vec2['aps'] and value != vec2['aps'][p]:
                        match = False
                if match:
                    removal.append(n)
                    removal.append(m)
        return removal

    # build_database.py:320-336
    @staticmethod
    def pack_ramified_primes(genus):
        return ",".join(map(str, genus.ramified_primes()))

    @staticmethod
    def unpack_ramified_primes(s):
        return list(map(int, s.split(",")))

    @staticmethod
    def pack_vector(vec):
        return ",".join(map(str, list(vec)))

    @staticmethod
    def unpack_vector(s):
        return [int(x) for x in s.split(",")]
