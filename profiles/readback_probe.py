"""Read-back of one resident T_p (C4, p = 9973, all 8 conductors) into pinned host memory, alone on the
machine: milliseconds for the plain int32 DMA and for the narrow (uint16 + int8, host-widened) path."""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ternary_birch_b200 import Genus, load_genus_table  # noqa: E402

table = load_genus_table(ROOT / "tests" / "golden" / "genus_1009091.npz")
for narrow in (False, True):
    if narrow:
        os.environ.pop("TBIRCH_NO_NARROW_COPY", None)
    else:
        os.environ["TBIRCH_NO_NARROW_COPY"] = "1"
    g = Genus(table, precise=False)
    res = g.hecke_device(9973)
    C = res.num_conductors()
    bufs = []
    for k in range(C):
        rows, _, nnz = res.shape(k)
        bufs.append((torch.empty(nnz, dtype=torch.int32).pin_memory(), torch.empty(nnz, dtype=torch.int32).pin_memory(),
                     torch.empty(rows + 1, dtype=torch.int32).pin_memory()))
    side = torch.cuda.Stream()
    best = 1e9
    times = []
    for rep in range(int(os.environ.get('REPS', 4))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(C):
            d, i, ptr = bufs[k]
            res.copy_sparse_to_async(k, d.data_ptr(), i.data_ptr(), ptr.data_ptr(), side.cuda_stream)
        side.synchronize()
        res.wait_copies()
        times.append((time.perf_counter() - t0) * 1e3)
        best = min(best, times[-1])
    pcie = sum(res.transfer_bytes(k) for k in range(C))
    total = sum(8 * res.shape(k)[2] + 4 * (res.shape(k)[0] + 1) for k in range(C))
    print(f"narrow={narrow} threads={os.environ.get('TBIRCH_HOST_THREADS', 'default')}: {best:.1f} ms, "
          f"{pcie / 1e9:.2f} GB over PCIe, {total / 1e9:.2f} GB delivered = {total / best / 1e6:.1f} GB/s; all reps: "
          + " ".join(f"{t:.0f}" for t in times))
    res.close()
    g.close()
