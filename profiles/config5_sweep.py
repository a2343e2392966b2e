#!/usr/bin/env python
"""BASELINE.json config 5 (sampled): prime sweep up to 10^5 on the mid-size genus of level
11*13*17*19*23 (N = 2016, 32 conductors) in precise (128-bit) mode.

p < 65536: T_p through hecke_matrix (every STRIDE-th good prime), matrices left on the device,
conductor-1 row sums checked on the device for every 8th sampled prime.
65536 < p < 10^5: eigenvalues() of the constant eigenvector (a_p = p + 1) -- the reference's Hecke
path truncates p to 16 bits, so parity there is only defined through eigenvalues.
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

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 16
t = load_genus_table(ROOT / "tests" / "golden" / "genus_1062347.npz")
g = Genus(t, precise=True, stream=torch.cuda.current_stream().cuda_stream)
limit = 100000
sieve = np.ones(limit + 1, dtype=bool)
sieve[:2] = False
for i in range(2, 317):
    if sieve[i]:
        sieve[i * i::i] = False
primes = [int(p) for p in np.nonzero(sieve)[0] if 1062347 % p]
small = [p for p in primes if p < 65536][::stride]
dev = torch.device("cuda")
g.hecke_device(small[-1]).close()
torch.cuda.synchronize()
t0 = time.perf_counter()
neighbors, kernel_ms, checked = 0, 0.0, 0
per_band = {}
for i, p in enumerate(small):
    res = g.hecke_device(p)
    a, b, n = g.last_kernel_time()
    kernel_ms += a + b
    neighbors += n
    band = "p<1e3" if p < 1000 else "p<1e4" if p < 10000 else "p<65536"
    acc = per_band.setdefault(band, [0, 0.0])
    acc[0] += n
    acc[1] += (a + b) * 1e-3
    if i % 8 == 0:
        rows, dim, nnz = res.shape(0)
        data = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        indptr = torch.empty(rows + 1, dtype=torch.int32, device=dev)
        res.copy_sparse_to(0, data.data_ptr(), indices.data_ptr(), indptr.data_ptr())
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(data[:nnz].to(torch.int64), 0)])
        sums = csum[indptr[1:].long()] - csum[indptr[:-1].long()]
        assert bool((sums == p + 1).all()), p
        checked += 1
    res.close()
torch.cuda.synchronize()
wall = time.perf_counter() - t0

b = BirchGenus(t)
b.add_eigenvector(dict(vector=[1] * t.N, conductor=1))
big = [p for p in primes if p > 65536][::stride * 4]
t1 = time.perf_counter()
for p in big:
    assert b.compute_eigenvalues(p, precise=True) == [p + 1], p
wall_big = time.perf_counter() - t1
print(json.dumps(dict(config="C5 (sampled): level 1062347, precise, every %d-th good prime" % stride, hecke_primes=len(small),
                      neighbors=neighbors, wall_s=wall, kernel_s=kernel_ms * 1e-3, neighbors_per_s=neighbors / wall,
                      kernel_neighbors_per_s={k: v[0] / v[1] for k, v in per_band.items()}, row_sum_checks=checked,
                      eigen_primes_above_2_16=len(big), eigen_wall_s=wall_big,
                      full_sweep_estimate_s=wall * stride,
                      reference_gmp_estimate_core_days=neighbors * stride / 0.17e6 / 86400)))
