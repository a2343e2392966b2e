"""Wall time of BirchGenus(level) -- form construction + genus enumeration on the GPU --
for the levels of the configs, and of the config-3 flow end to end without Sage."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ternary_birch_b200 import BirchGenus  # noqa: E402

for level in (11, 85085, 1062347, 1009091):
    t0 = time.perf_counter()
    b = BirchGenus(level, seed=1)
    t1 = time.perf_counter()
    print(f"BirchGenus({level}): N = {b.table.N}, {len(b.dims)} conductors, {t1 - t0:.3f} s")
    if level == 85085:
        t0 = time.perf_counter()
        vecs = b.rational_eigenvectors()
        t1 = time.perf_counter()
        b.compute_eigenvalues_upto(1000)
        t2 = time.perf_counter()
        print(f"  config 3: rational_eigenvectors {t1 - t0:.3f} s ({len(vecs)} vectors), "
              f"compute_eigenvalues_upto(1000) {t2 - t1:.3f} s")
