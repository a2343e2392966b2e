import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = Path(__file__).resolve().parent / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: minutes-long parity case")


def pytest_sessionstart(session):
    # tests are allowed to use the checkers; make sure they are built
    # (the C port always; oracle/_ref only where /root/reference exists)
    from oracle import pyoracle
    pyoracle.build()
    # the product library: (re)built in-tree when missing or stale (nvcc cross-compiles without a GPU)
    import __graft_entry__ as ge
    ge.build()


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def primes_upto(n):
    sieve = np.ones(n + 1, dtype=bool)
    sieve[:2] = False
    for i in range(2, int(n ** 0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = False
    return [int(p) for p in np.nonzero(sieve)[0]]


def csr_digest(conductor, dim, data, indices, indptr) -> dict:
    h = hashlib.sha256()
    for a in (data, indices, indptr):
        h.update(np.ascontiguousarray(a, dtype="<i4").tobytes())
    return dict(conductor=int(conductor), dim=int(dim), nnz=int(np.asarray(data).size), sha256=h.hexdigest())


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def tables():
    from ternary_birch_b200 import load_genus_table
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_genus_table(GOLDEN / f"genus_{name}.npz")
        return cache[name]
    return get


@pytest.fixture(scope="session")
def hecke_small():
    return dict(np.load(GOLDEN / "hecke_small.npz"))


@pytest.fixture(scope="session")
def records_small():
    return dict(np.load(GOLDEN / "records_small.npz"))


@pytest.fixture(scope="session")
def hecke_digests():
    d = json.load(open(GOLDEN / "hecke_digests.json"))
    slow = GOLDEN / "hecke_digests_slow.json"
    if slow.exists():
        d.update(json.load(open(slow)))
    return d


@pytest.fixture(scope="session")
def port_oracle(tables):
    """name -> PortOracle (the plain-C restatement) on that genus table."""
    from oracle.pyoracle import GenusArrays, PortOracle
    cache = {}

    def get(name):
        if name not in cache:
            t = tables(name)
            ga = GenusArrays(t.N, t.K, t.seed, t.disc, t.primes, t.conductors, t.dims, t.q, t.s, t.sinv,
                             t.scalar, t.parent, t.p, t.num_auts, t.lut)
            cache[name] = PortOracle(ga)
        return cache[name]
    return get
