import sys, time, os
sys.path.insert(0, '.')
import torch
from ternary_birch_b200 import Genus, load_genus_table
t = load_genus_table('tests/golden/genus_1009091.npz')
g = Genus(t, precise=False, stream=torch.cuda.current_stream().cuda_stream)
g.hecke_device(9973).close()
torch.cuda.synchronize()
for label, primes in (("same p=9973 x40", [9973] * 40), ("ascending 40 primes near 5000", [p for p in range(5000, 5400) if all(p % d for d in range(2, 80))][:40])):
    t0 = time.perf_counter(); k = 0.0
    for p in primes:
        res = g.hecke_device(p)
        a, b, n = g.last_kernel_time(); k += a + b
        res.close()
    torch.cuda.synchronize()
    w = time.perf_counter() - t0
    print(f"{label}: wall {w*1e3/len(primes):.2f} ms/call, kernels {k/len(primes):.2f} ms/call, overhead {(w*1e3-k)/len(primes):.2f} ms/call")
