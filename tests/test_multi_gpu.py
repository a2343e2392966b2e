"""GPU, world_size 2 over NCCL: each rank computes its shard of a T_p on its own GPU through the C ABI
(`tb_hecke_compute_rows`), the shards meet on rank 0 through the final integer gather
(`tb_hecke_pack` -> grouped send/recv -> `tb_unpack_rows`, ternary_birch_b200/shard.py) and the
reassembled matrices are compared with the oracle entry for entry.  Needs two GPUs: run it with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu` (skipped on one)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = Path(__file__).resolve().parent / "golden"

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, result_path, name, p, precise, narrow):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if narrow is not None:
        os.environ["TBIRCH_NARROW_MIN_NNZ"] = "1"
        os.environ["TBIRCH_FORCE_NARROW_COPY" if narrow else "TBIRCH_NO_NARROW_COPY"] = "1"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle.pyoracle import GenusArrays, PortOracle
    from ternary_birch_b200 import Genus, load_genus_table, shard
    t = load_genus_table(GOLDEN / f"genus_{name}.npz")
    g = Genus(t, precise=precise, stream=torch.cuda.current_stream().cuda_stream)
    # (1) one T_p split by rows: rank r owns a contiguous block of representatives
    begin, end = shard.row_shards(t.N, world)[rank]
    res = g.hecke_device(p, begin, end)
    shard.warm_gather(dev)
    got = shard.gather_hecke_device(res, dev)
    ok, why = True, ""
    if rank == 0:
        parts, info = got
        host_parts = [{c: tuple(a.cpu().numpy() for a in trio) for c, trio in part.items()} for part in parts]
        whole = shard.concat_row_shards(host_parts)
        oracle = PortOracle(GenusArrays(t.N, t.K, t.seed, t.disc, t.primes, t.conductors, t.dims, t.q, t.s, t.sinv,
                                        t.scalar, t.parent, t.p, t.num_auts, t.lut))
        for m in oracle.hecke(p, precise):
            d, i, ptr = whole[m["conductor"]]
            if not (np.array_equal(d, m["data"]) and np.array_equal(i, m["indices"]) and np.array_equal(ptr, m["indptr"])):
                ok, why = False, f"row shards: conductor {m['conductor']}"
        # columns fit 16 bits and p + 1 < 128 neighbors per row: the 3-byte form travels whatever the row kernel wrote
        if info["entry_bytes"] != [3] * world:
            ok, why = False, f"expected the 3-byte form on every rank, got {info['entry_bytes']}"
    res.close()
    # (2) prime axis: rank r computes a different prime; rank 0 receives every T_p whole
    primes = shard.prime_shards([p, 29, 31, 37][:world], world)          # none divides a golden level
    mine = primes[rank][0]
    res = g.hecke_device(mine)
    got = shard.gather_hecke_device(res, dev)
    if rank == 0:
        parts, _ = got
        for r, part in enumerate(parts):
            for m in oracle.hecke(primes[r][0], precise):
                d, i, ptr = (a.cpu().numpy() for a in part[m["conductor"]])
                if not (np.array_equal(d, m["data"]) and np.array_equal(i, m["indices"]) and np.array_equal(ptr, m["indptr"])):
                    ok, why = False, f"prime shards: rank {r} conductor {m['conductor']}"
        Path(result_path).write_text("ok" if ok else "mismatch: " + why)
    res.close()
    g.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,p,precise,narrow", [("85085", 101, False, True), ("85085", 101, True, False),
                                                   ("1062347", 101, False, None)])
def test_row_and_prime_shards_gather_world2(tmp_path, name, p, precise, narrow):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, port, str(result), name, p, precise, narrow), nprocs=2, join=True)
    assert result.read_text() == "ok"
