#!/usr/bin/env python
"""Latency of the small named calls of BASELINE.json configs 1-3 (wall clock, host buffers out).

  C1  BirchGenus(11): T_p sparse for all 24 good p < 100, int64
  C2  BirchGenus(11*13*17*19*23): hecke_matrix(101, *) sparse + dense, int64 and precise
  C3  BirchGenus(5*7*11*13*17): eigenvalues of the 9 rational eigenvectors for all 163 good p <= 1000
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ternary_birch_b200 import BirchGenus, Genus, load_genus_table  # noqa: E402

G = ROOT / "tests" / "golden"


def timed(fn, repeat=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(repeat):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


out = {}
t = load_genus_table(G / "genus_11.npz")
g = Genus(t, precise=False)
primes = [p for p in range(2, 100) if all(p % d for d in range(2, p)) and p != 11]
out["C1_24_primes_sparse_int64_ms"] = 1e3 * timed(lambda: [g.hecke_matrix_sparse(p) for p in primes])

t = load_genus_table(G / "genus_1062347.npz")
for precise in (False, True):
    g = Genus(t, precise=precise)
    tag = "precise" if precise else "int64"
    out[f"C2_T101_sparse_{tag}_ms"] = 1e3 * timed(lambda: g.hecke_matrix_sparse(101))
    out[f"C2_T101_dense_{tag}_ms"] = 1e3 * timed(lambda: g.hecke_matrix_dense(101), repeat=2)
    res = g.hecke_device(101)
    res.close()
    out[f"C2_T101_device_{tag}_ms"] = 1e3 * timed(lambda: g.hecke_device(101).close())
    out[f"C2_T101_kernels_{tag}_ms"] = sum(g.last_kernel_time()[:2])

gold = json.load(open(G / "eigen_85085.json"))
for precise in (False, True):
    b = BirchGenus(load_genus_table(G / "genus_85085.npz"))
    for v in gold["vectors"]:
        b.add_eigenvector(v)
    b.compute_eigenvalues(3, precise=precise)
    t0 = time.perf_counter()
    b.compute_eigenvalues_upto(1000, precise=precise, force=True)
    out[f"C3_eigenvalues_upto_1000_{'precise' if precise else 'int64'}_ms"] = 1e3 * (time.perf_counter() - t0)
print(json.dumps(out, indent=1))

# C4-size call into fresh numpy arrays (pageable memory): the whole hecke_matrix_sparse(9973)
t = load_genus_table(G / "genus_1009091.npz")
g = Genus(t, precise=False)
g.hecke_matrix_sparse(107)
t0 = time.perf_counter()
m = g.hecke_matrix_sparse(9973)
print(json.dumps({"C4_T9973_sparse_into_numpy_ms": 1e3 * (time.perf_counter() - t0),
                  "C4_T9973_bytes": int(sum(a.nbytes for trio in m.values() for a in trio))}))
