"""One GPU elimination of T_2 - e (level-1009091 genus, conductor 1, N = 10316) modulo 2^29 - 3:
the unit of work of the rational eigenvector search, for ncu launch lists."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ternary_birch_b200 import BirchGenus, eigen, load_genus_table  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "1009091"
b = BirchGenus(load_genus_table(ROOT / "tests" / "golden" / f"genus_{name}.npz"))
A = b.hecke_matrix(2, 1, sparse=True, precise=True)
n = A.shape[0]
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    free, basis, ms = eigen.kernel_mod_csr(A.data, A.indices, A.indptr, n, 1, 536870909)
    print(f"N={n} kernel_dim={len(free)} gpu_ms={ms:.2f} modmul/s={n ** 3 / (ms * 1e-3):.3e}")
