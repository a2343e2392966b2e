#!/usr/bin/env python
"""BASELINE.json config 4 in full on one GPU: T_p for EVERY good prime p <= 10^4 on the genus of
level 97*101*103 (N = 10316, 8 conductors), int64, matrices left on the device.

For every prime the 8 CSR matrices are computed; every 25th prime the conductor-1 matrix is
checked on the device (each row sums to p + 1).  Prints one JSON summary.
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from ternary_birch_b200 import Genus, load_genus_table  # noqa: E402

limit = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
t = load_genus_table(ROOT / "tests" / "golden" / "genus_1009091.npz")
g = Genus(t, precise=False, stream=torch.cuda.current_stream().cuda_stream)
sieve = [True] * (limit + 1)
primes = []
for p in range(2, limit + 1):
    if sieve[p]:
        primes.append(p)
        for m in range(p * p, limit + 1, p):
            sieve[m] = False
primes = [p for p in primes if 1009091 % p]
dev = torch.device("cuda")
g.hecke_device(primes[-1]).close()          # warm-up (pool, kernels)
torch.cuda.synchronize()
t0 = time.perf_counter()
neighbors = 0
kernel_ms = 0.0
checked = 0
nnz_total = 0
for i, p in enumerate(primes):
    res = g.hecke_device(p)
    a, b, n = g.last_kernel_time()
    kernel_ms += a + b
    neighbors += n
    nnz_total += res.total_nnz()
    if i % 25 == 0:
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
print(json.dumps(dict(config="C4: level 1009091, all good primes <= %d" % limit, primes=len(primes), neighbors=neighbors,
                      wall_s=wall, kernel_s=kernel_ms * 1e-3, neighbors_per_s=neighbors / wall,
                      csr_entries=nnz_total, row_sum_checks=checked,
                      reference_estimate_core_hours=neighbors / 2.0e6 / 3600)))
