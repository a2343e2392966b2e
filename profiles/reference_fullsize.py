"""Pin the sampled reference rate: the UNMODIFIED reference binary (oracle/_ref/birch_ref) at FULL size
on this box -- one OS process per prime among the largest primes below 10^4 (the bench workload's
range), as many processes as host cores (capped), Genus::hecke_matrix_sparse(p) in int64 on the C4
genus -- next to the same measurement on the small primes bench.py samples.  ~60-100 s of CPU."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

cores = os.cpu_count() or 1
# each full-size process holds its T_p as growing std::vector<int>s (3.6 GB of CSR, ~10 GB at the peak of a
# reallocation): as many processes as the host's free memory allows, 16 at most
avail_gb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")) / 1e6
n = max(1, min(cores, 16, int(avail_gb // 12)))
big = [9973, 9967, 9949, 9941, 9931, 9929, 9923, 9907, 9901, 9887, 9883, 9871, 9859, 9857, 9851, 9839][:n]
notes = []
while True:
    t0 = time.time()
    try:
        neighbors, slowest, kind = bench.reference_sample(big, timeout=3000)
        break
    except RuntimeError as e:      # e.g. std::bad_alloc with too many 10 GB processes at once: halve and retry
        notes.append(f"{n} processes: {e}")
        if n == 1:
            raise
        n //= 2
        big = big[:n]
wall = time.time() - t0
small = bench.REF_SAMPLE_PRIMES[:n]
n2, s2, _ = bench.reference_sample(small)
rec = dict(kind=kind, cores_used=n, host_cores=cores,
           full=dict(primes=big, neighbors=neighbors, slowest_call_s=slowest, wall_s=wall,
                     neighbors_per_s=neighbors / slowest, neighbors_per_s_per_core=neighbors / slowest / n),
           sampled=dict(primes=small, neighbors=n2, slowest_call_s=s2, neighbors_per_s=n2 / s2,
                        neighbors_per_s_per_core=n2 / s2 / n),
           neighbors_per_s_per_core=neighbors / slowest / n,
           what=f"unmodified reference, {n} processes x one prime each of {big[0]}..{big[-1]}, int64, "
                f"{slowest:.1f} s for the slowest call")
if notes:
    rec["retries"] = notes
rec["sampled_over_full"] = rec["sampled"]["neighbors_per_s_per_core"] / rec["full"]["neighbors_per_s_per_core"]
print(json.dumps(rec))
