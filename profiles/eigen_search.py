"""Timing of the rational eigenvector search on one B200.

usage: python profiles/eigen_search.py [level-table ...]   (default: 85085 1009091)
Prints one JSON line per genus: eliminations, GPU ms, wall seconds, eigenvectors found;
and the device time of single eliminations of T_2 - e at that size."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ternary_birch_b200 import BirchGenus, eigen, load_genus_table  # noqa: E402


def main():
    names = sys.argv[1:] or ["85085", "1009091"]
    for name in names:
        table = load_genus_table(ROOT / "tests" / "golden" / f"genus_{name}.npz")
        b = BirchGenus(table)
        p = b.next_good_prime(1)
        A = b.hecke_matrix(p, 1, sparse=True, precise=True)
        n = A.shape[0]
        single = {}
        for e in (0, 1):
            free, basis, ms = eigen.kernel_mod_csr(A.data, A.indices, A.indptr, n, e, 536870909)
            single[e] = dict(kernel_dim=len(free), gpu_ms=round(ms, 2),
                             modmul_per_s=round(float(n) ** 3 / (ms * 1e-3), 1) if ms > 0 else None)
        t0 = time.time()
        lines = []
        vecs = b.rational_eigenvectors(precise=True, log=lines.append)
        wall = time.time() - t0
        st = b.eigenvector_search_stats
        print(json.dumps(dict(level=int(table.level), N=int(table.N), p=p, single_elimination=single,
                              kernels=st["kernels"], eliminations=st["primes"], gpu_ms=round(st["gpu_ms"], 1),
                              wall_s=round(wall, 2), eigenvectors=len(vecs),
                              by_conductor={int(c): sum(1 for v in vecs if v["conductor"] == c) for c in b.dims})),
              flush=True)
        for ln in lines[:400]:
            print("  " + ln)


if __name__ == "__main__":
    main()
