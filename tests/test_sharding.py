"""CPU, world_size 2 over gloo: the multi-GPU partition and the final integer gather of
Hecke rows (ternary_birch_b200/shard.py).  Per-rank rows come from the oracle port (no GPU
here); the product code under test is the partitioning, packing, gather and reassembly."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = Path(__file__).resolve().parent / "golden"


def _rows_of(mats, lut, begin, end):
    """Slice the full CSR matrices to the rows of representatives [begin, end)."""
    out = {}
    for k, m in enumerate(mats):
        inc = [n for n in range(begin, end) if lut[k][n] != -1]
        if inc:
            r0, r1 = int(lut[k][inc[0]]), int(lut[k][inc[-1]]) + 1
        else:
            r0 = r1 = 0
        lo, hi = int(m["indptr"][r0]), int(m["indptr"][r1])
        out[m["conductor"]] = (m["data"][lo:hi], m["indices"][lo:hi], (m["indptr"][r0:r1 + 1] - lo).astype(np.int32))
    return out


def _worker(rank, world, port, result_path):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import GenusArrays, PortOracle
    from ternary_birch_b200 import load_genus_table, shard
    t = load_genus_table(GOLDEN / "genus_85085.npz")
    port_oracle = PortOracle(GenusArrays(t.N, t.K, t.seed, t.disc, t.primes, t.conductors, t.dims, t.q, t.s, t.sinv,
                                         t.scalar, t.parent, t.p, t.num_auts, t.lut))
    full = port_oracle.hecke(19, False)
    begin, end = shard.row_shards(t.N, world)[rank]
    local = _rows_of(full, t.lut, begin, end)
    parts = shard.gather_csr(local)
    ok = True
    if rank == 0:
        whole = shard.concat_row_shards(parts)
        for m in full:
            d, i, p = whole[m["conductor"]]
            ok &= bool(np.array_equal(d, m["data"]) and np.array_equal(i, m["indices"]) and np.array_equal(p, m["indptr"]))
        # prime-axis sharding: every prime lands on exactly one rank, largest first
        dealt = shard.prime_shards([2, 3, 19, 23, 29, 31, 37], world)
        ok &= sorted(sum(dealt, [])) == [2, 3, 19, 23, 29, 31, 37] and dealt[0][0] == 37
        Path(result_path).write_text("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_row_shards_gather_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, port, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"


def test_shard_helpers():
    from ternary_birch_b200 import shard
    assert shard.row_shards(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard.row_shards(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    local = {1: (np.array([3, 4], np.int32), np.array([0, 1], np.int32), np.array([0, 2], np.int32)),
             11: (np.zeros(0, np.int32), np.zeros(0, np.int32), np.array([0], np.int32))}
    h, p = shard.pack_csr(local)
    back = shard.unpack_csr(h, p)
    assert sorted(back) == [1, 11] and back[1][0].tolist() == [3, 4] and back[11][2].tolist() == [0]
