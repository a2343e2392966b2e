#!/usr/bin/env python
"""bench.py -- neighbors/sec of the p-neighbor Hecke hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Workload ("C4", BASELINE.json configs[3], the configuration the 1/2/4/8-GPU metric is
quoted on): the genus of level 97*101*103 = 1009091 (N = 10316 representatives, K = 3,
8 conductors, seed 1; committed fixture tests/golden/genus_1009091.npz) and T_p for a
prime just below 10^4, int64 mode, CSR output for all conductors.  One STEP is one
T_p: (p+1) * N neighbors enumerated, built, Eisenstein-reduced, looked up in the genus
hash table, given their spinor character and accumulated into Hecke rows.  Under
torchrun rank i takes the i-th largest prime below 10^4 (weak scaling: one T_p per
rank per step, no data-path collective; a final integer gather of the rows to rank 0
runs once after the timed region).

One JSON line is printed by rank 0.  `value` has the genus table resident in HBM;
`e2e` re-sends the table and the per-prime inputs host->device every step and reads
every CSR array back to pinned host memory.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

LEVEL = 1009091
TABLE = ROOT / "tests" / "golden" / f"genus_{LEVEL}.npz"
PRIMES_DESC = [9973, 9967, 9949, 9941, 9931, 9929, 9923, 9907]   # largest primes below 10^4
REF_SAMPLE_PRIMES = [1009, 1013, 1019, 1021, 1031, 1033, 1039, 1049, 1051, 1061, 1063, 1069, 1087, 1091, 1093, 1097,
                     1103, 1109, 1117, 1123, 1129, 1151, 1153, 1163, 1171, 1181, 1187, 1193, 1201, 1213, 1217, 1223,
                     1229, 1231, 1237, 1249, 1259, 1277, 1279, 1283, 1289, 1291, 1297, 1301, 1303, 1307, 1319, 1321,
                     1327, 1361, 1367, 1373, 1381, 1399, 1409, 1423, 1427, 1429, 1433, 1439, 1447, 1451, 1453, 1459]
METRIC = "neighbors_per_sec"
UNIT = "neighbors/s"


def workload_name(p):
    return (f"C4: genus of level 97*101*103=1009091 (N=10316, K=3, seed 1), T_p all 8 conductors, "
            f"p={p} (rank i: i-th largest prime < 10^4), int64, CSR")


def mean_reduce_iters(p: int) -> float:
    # measured 4.86 / 7.04 / 9.33 / 9.99 at p = 101 / 997 / 10007 / 65521
    return 0.6 * math.log2(p) + 1.0


def int32_ops_per_neighbor(p: int, K: int) -> float:
    """algorithmic work per neighbor in 32-bit integer instructions."""
    it = mean_reduce_iters(p)
    mul64 = 101 + 12 * it
    add64 = 107 + 87 * it + 4 * (1 << K)
    div64 = 10 + K + 0.33 * it
    return 4 * mul64 + 2 * add64 + 20 * div64 + 70


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, cell in zip(names, r[5:9]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None),
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref = unmodified reference sources compiled here)
# ---------------------------------------------------------------------------
def reference_sample(primes, precise=False, timeout=1200):
    """One OS process per prime (the reference's own parallel model,
    scripts/generate_eigenvalues.py:205), each timing Genus::hecke_matrix_sparse(p) on
    the workload's genus.  Returns (total neighbors, wall seconds of the slowest call, kind)."""
    from oracle import pyoracle          # the cpu_baseline / --impl reference leg may execute oracle/
    if pyoracle.RefBinary.available():
        kind = "reference"
        mode = "z" if precise else "z64"
        procs = [subprocess.Popen([str(pyoracle.REF_BIN), "--level", str(LEVEL), "--seed", "1", "time", str(p),
                                   mode, "sparse", "1"], stdout=subprocess.PIPE, text=True) for p in primes]
        total, slowest = 0, 0.0
        for pr in procs:
            out, _ = pr.communicate(timeout=timeout)
            lines = [l for l in out.splitlines() if l.startswith("{")]
            rec = json.loads(lines[-1]) if lines else {"error": "no output", "what": f"exit code {pr.returncode}"}
            if "neighbors" not in rec:
                for other in procs:
                    if other.poll() is None:
                        other.kill()
                raise RuntimeError(f"reference binary failed: {rec}")
            total += rec["neighbors"]
            slowest = max(slowest, rec["seconds"])
        return total, slowest, kind
    # oracle/_ref did not travel: time the plain-C port instead, one process per prime
    kind = "port"
    code = ("import sys,time;sys.path.insert(0,%r);import numpy as np;"
            "from oracle.pyoracle import PortOracle,load_genus;from ternary_birch_b200 import load_genus_table;"
            "from oracle.pyoracle import GenusArrays as G;t=load_genus_table(%r);"
            "o=PortOracle(G(t.N,t.K,t.seed,t.disc,t.primes,t.conductors,t.dims,t.q,t.s,t.sinv,t.scalar,t.parent,t.p,t.num_auts,t.lut));"
            "p=int(sys.argv[1]);t0=time.perf_counter();o.hecke(p,%s);print(t.N*(p+1),time.perf_counter()-t0)"
            % (str(ROOT), str(TABLE), "True" if precise else "False"))
    procs = [subprocess.Popen([sys.executable, "-c", code, str(p)], stdout=subprocess.PIPE, text=True) for p in primes]
    total, slowest = 0, 0.0
    for pr in procs:
        out, _ = pr.communicate(timeout=timeout)
        n, s = out.split()[-2:]
        total += int(n)
        slowest = max(slowest, float(s))
    return total, slowest, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    primes = REF_SAMPLE_PRIMES[:min(cores, len(REF_SAMPLE_PRIMES))]
    for _ in range(max(args.warmup, 0)):
        reference_sample(primes[:1] if args.warmup > 1 else primes)
    total_n, total_s, kind = 0, 0.0, "reference"
    for _ in range(args.steps):
        n, s, kind = reference_sample(primes)
        total_n += n
        total_s += s
    value = total_n / total_s
    sample = (f"per step: Genus::hecke_matrix_sparse(p) int64 on the same genus for the {len(primes)} primes "
              f"{primes[0]}..{primes[-1]}, one OS process per prime on {cores} host cores "
              f"(full-size p~10^4 costs ~50 s per prime per core)")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total_s / max(args.steps, 1), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="int64", data="synthetic",
                config=dict(workload=workload_name(PRIMES_DESC[0]), sample=sample),
                cpu_baseline=dict(value=value, unit=UNIT, cores=len(primes), kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)
    return 0


def host_ingest_summary(per_rank, csr_bytes_per_neighbor, e2e_value):
    """The two ceilings of delivering int32 CSR into this host's memory, measured with all ranks
    active at once right after the end-to-end loop: plain DMA into pinned memory, and the host
    threads' streaming stores (the widening path).  End to end cannot beat the larger one."""
    dma = [r["ingest"]["dma_gbs"] for r in per_rank]
    wid = [r["ingest"]["widen_gbs"] for r in per_rank]
    c_dma = sum(dma) * 1e9 / csr_bytes_per_neighbor
    c_wid = sum(wid) * 1e9 / csr_bytes_per_neighbor
    # weak scaling is timed as the max over ranks: with plain DMA the slowest link sets the step
    c_dma_slowest = len(dma) * min(dma) * 1e9 / csr_bytes_per_neighbor
    return dict(csr_bytes_per_neighbor=round(csr_bytes_per_neighbor, 2), dma_gbs_per_rank=dma, dma_gbs_total=round(sum(dma), 1),
                widen_gbs_per_rank=wid, widen_gbs_total=round(sum(wid), 1), widen_threads_per_rank=per_rank[0]["ingest"]["widen_threads"],
                ceiling_dma_neighbors_per_s=c_dma, ceiling_dma_slowest_rank_neighbors_per_s=c_dma_slowest,
                ceiling_widen_neighbors_per_s=c_wid, e2e_frac_of_ceiling=e2e_value / max(c_dma, c_wid),
                note="all ranks concurrently; see profiles/r2a_d2h_probe_n4.md")


def fullsize_note(cores, sampled_total):
    """The sampled reference rate next to the one full-size timing committed under profiles/."""
    f = ROOT / "profiles" / "r2_reference_fullsize.json"
    if not f.exists():
        return ""
    try:
        rec = json.load(open(f))
        return (f"; sampled {sampled_total / cores / 1e6:.2f} M neighbors/s/core here vs {rec['neighbors_per_s_per_core'] / 1e6:.2f} M/s/core "
                f"at full size ({rec['what']})")
    except Exception:
        return ""


def bind_to_gpu_numa_node(torch, local_rank: int):
    """Best effort: run this rank (and so its pinned-memory allocations, first touch) on the CPUs of
    the NUMA node its GPU hangs off, so the end-to-end copies of 8 ranks do not cross sockets."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precise", action="store_true", help="the reference's arbitrary-precision mode (Genus<Z>) instead of int64")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from ternary_birch_b200 import Genus, cabi, load_genus_table, shard

    table = load_genus_table(TABLE)
    p = PRIMES_DESC[rank % len(PRIMES_DESC)]
    N, K = table.N, table.K
    neighbors_per_step = N * (p + 1)
    stream = torch.cuda.current_stream()
    g = Genus(table, precise=args.precise, stream=stream.cuda_stream)
    L = cabi.lib()
    launches0 = L.tb_launch_count()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident arm (`value`) -------------------------------------
    def step_resident():
        res = g.hecke_device(p)
        return res

    for _ in range(args.warmup):
        step_resident().close()
    sampler = ClockSampler(local_rank)
    sync_all()
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kernel_ms, rows_ms = [], []
    launches_before = L.tb_launch_count()
    last = None
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                      # L2 flush between timed iterations (untimed)
        starts[i].record(stream)
        if last is not None:
            last.close()
        last = step_resident()
        stops[i].record(stream)
        a, b, _ = g.last_kernel_time()
        kernel_ms.append(a)
        rows_ms.append(b)
    sync_all()
    launches_timed = L.tb_launch_count() - launches_before
    clocks = sampler.stop() if rank == 0 else None
    my_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    t = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    total_neighbors = 0
    for r in range(world):
        total_neighbors += N * (PRIMES_DESC[r % len(PRIMES_DESC)] + 1) * args.steps
    value = total_neighbors / (total_ms * 1e-3)

    # ---- end-to-end arm (`e2e`): host buffers in, host CSR out, every step ----
    C = 1 << K
    shapes = [last.shape(k) for k in range(C)]
    pinned = []
    for rows, _, nnz in shapes:
        pinned.append((torch.empty(max(nnz, 1) + 4096, dtype=torch.int32).pin_memory(),
                       torch.empty(max(nnz, 1) + 4096, dtype=torch.int32).pin_memory(),
                       torch.empty(rows + 1, dtype=torch.int32).pin_memory()))
    h2d_bytes = g.table_bytes() + 12 * N
    result_bytes = sum(4 * (2 * nnz + rows + 1) for rows, _, nnz in shapes)       # int32 CSR the caller receives
    d2h_bytes = sum(last.transfer_bytes(k) for k in range(C))                     # bytes that cross PCIe for it
    last.close()

    # Two pinned result sets and a copy stream: the read-back of step i overlaps the compute of
    # step i+1 (each step still moves its own inputs in and its own result out inside the timed region).
    pinned = [pinned, [tuple(t.clone().pin_memory() for t in trio) for trio in pinned]]
    copy_stream = torch.cuda.Stream(device=dev)
    pending = [None, None]                               # (HeckeResult, event) per buffer set

    def drain(slot):
        if pending[slot] is not None:
            res, ev = pending[slot]
            ev.synchronize()
            res.wait_copies()                            # host-widened read-back runs off-stream
            res.close()
            pending[slot] = None

    def step_e2e(i):
        slot = i & 1
        drain(slot)                                      # buffer set free again
        g.upload()                                       # genus table + per-prime inputs: host -> device
        res = g.hecke_device(p)                          # returns when the T_p is complete on the device
        copy_stream.wait_stream(stream)
        for k in range(C):
            d, i_, ptr = pinned[slot][k]
            rows, _, nnz = res.shape(k)
            assert nnz <= d.numel()
            res.copy_sparse_to_async(k, d.data_ptr(), i_.data_ptr(), ptr.data_ptr(), copy_stream.cuda_stream)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        pending[slot] = (res, ev)

    # warm-up: W untimed steps, at least eight: both pinned result sets must have been written by the host
    # once (the first CPU store to a page of a fresh cudaHostAlloc mapping takes a minor fault), and the library
    # measures one read-back in each format before it settles on one (tb_readback_stats)
    # (with several ranks per host the per-rank format choices settle against each other: at 8 ranks the
    #  measured steps were 340 ms for the first ten and 162 ms from then on, profiles/README.md r2y)
    e2e_warmup = max(8 if world == 1 else 24, args.warmup)
    for i in range(e2e_warmup):
        step_e2e(i)
    drain(0)
    drain(1)
    sync_all()
    e_start = torch.cuda.Event(enable_timing=True)
    e_stop = torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e_start.record(stream)
    step_marks = [wall0]
    for i in range(args.steps):
        step_e2e(i)
        step_marks.append(time.perf_counter())
    drain(0)
    drain(1)
    step_marks.append(time.perf_counter())
    e_stop.record(stream)
    sync_all()
    wall_e2e = time.perf_counter() - wall0
    e_ms = max(e_start.elapsed_time(e_stop), wall_e2e * 1e3)
    t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_neighbors / (float(t.item()) * 1e-3)
    my_steps = [round((b - a) * 1e3, 1) for a, b in zip(step_marks, step_marks[1:])]
    pol = [ctypes.c_double(), ctypes.c_double(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()]
    cabi.check(L.tb_readback_stats(*[ctypes.byref(x) for x in pol]))
    my_policy = dict(rank=rank, format="narrow" if pol[4].value else "plain", plain_gbs=round(pol[0].value, 1),
                     narrow_gbs=round(pol[1].value, 1), plain_samples=pol[2].value, narrow_samples=pol[3].value)

    # ---- what can this host ingest?  All ranks at once: plain DMA into pinned memory, then the
    # widening threads alone (profiles/r2a_d2h_probe_n4.md explains why these are the two ceilings).
    probe_src = flush[:256 << 20]
    probe_dst = pinned[0][0][0].view(torch.uint8)
    probe_len = min(probe_src.numel(), probe_dst.numel()) // 4096 * 4096
    sync_all()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 12
    pe0.record(stream)
    for _ in range(reps):
        probe_dst[:probe_len].copy_(probe_src[:probe_len], non_blocking=True)
    pe1.record(stream)
    torch.cuda.synchronize()
    my_dma = reps * probe_len / (pe0.elapsed_time(pe1) * 1e-3) / 1e9
    sync_all()
    wg, wt, wi = ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
    cabi.check(L.tb_host_widen_probe(16 << 20, ctypes.byref(wg), ctypes.byref(wt), ctypes.byref(wi)))
    my_ingest = dict(rank=rank, dma_gbs=round(my_dma, 1), widen_gbs=round(wg.value, 1), widen_threads=wt.value)
    per_rank = [dict(steps=my_steps, policy=my_policy, ingest=my_ingest)]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, dict(steps=my_steps, policy=my_policy, ingest=my_ingest))

    # ---- final integer gather of the Hecke rows (once, outside the timed region) -------
    gather = None
    if world > 1:
        res = g.hecke_device(p)
        shard.warm_gather(dev)                       # NCCL point-to-point connection set-up, once
        gms = []
        for _ in range(2):
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            got = shard.gather_hecke_device(res, dev)
            torch.cuda.synchronize()
            gms.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor(gms, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            parts, info = got
            assert len(parts) == world
            for r, part in enumerate(parts):
                # every rank's rows must arrive intact: conductor 1 has no character, so its entries
                # count neighbors and sum to (p_r + 1) * N; every indptr ends at its nnz
                pr = PRIMES_DESC[r % len(PRIMES_DESC)]
                data, indices, indptr = part[1]
                assert int(data.sum(dtype=torch.int64).item()) == (pr + 1) * N, f"gathered rows of rank {r} are damaged"
                assert int(indptr[-1].item()) == data.numel() and int(indices.max().item()) < N
                for cond, (d_, i_, ptr_) in part.items():
                    assert int(ptr_[-1].item()) == d_.numel() == i_.numel(), (r, cond)
            gather = dict(ms=float(t[1].item()), first_call_ms=float(t[0].item()), bytes_over_nvlink=int(sum(info["bytes_per_rank"][1:])),
                          entry_bytes=info["entry_bytes"], verified_ranks=world,
                          how="tb_hecke_pack (3-byte form) -> one all_gather of sizes -> grouped send/recv to rank 0 "
                              "-> tb_unpack_rows (int32 CSR on rank 0's device)")
            del parts, got
        res.close()

    # ---- roofline of the dominant kernel (neighbors_kernel) ------------------
    peak = None
    if rank == 0:
        arms = (ctypes.c_double * 7)()
        cabi.check(L.tb_int_peak_detail(arms))
        peak = dict(imad_only=arms[0], lop3_only=arms[1], add_only=arms[2], imad_add=arms[3], imad_lop3=arms[4])
        fp64_arms = dict(dfma_only=arms[5], dfma_imad_lop3=arms[6])
    kms = sorted(kernel_ms)[len(kernel_ms) // 2]
    rms = sorted(rows_ms)[len(rows_ms) // 2]

    if rank == 0:
        w32 = int32_ops_per_neighbor(p, K)
        achieved = neighbors_per_step * w32 / (kms * 1e-3) / 1e12
        measured_peak = max(peak.values()) / 1e12
        sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        nominal_peak = torch.cuda.get_device_properties(dev).multi_processor_count * 4 * 32 * sm_mhz * 1e6 / 1e12
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak = json.load(open(peaks_file))["hbm_gbs"] if peaks_file.exists() else 6650.0
        # algorithmic DRAM-side bytes per neighbor: the 4-byte packed word written, plus the packed
        # word re-read 3x by the row kernels; the genus table (<3 MB) stays in L2
        hbm_achieved = neighbors_per_step * 4.0 / (kms * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "neighbors_kernel_traffic.json"
        if tf.exists():
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        roofline = dict(bound="int32-issue", achieved=achieved, peak=measured_peak, unit="Tint32op/s",
                        frac=achieved / measured_peak, frac_of_nominal_issue=achieved / nominal_peak,
                        nominal_issue_peak=nominal_peak, traffic=traffic,
                        kernel="neighbors_kernel<%s,PACKED>" % {"int64-wrap": "W64", "int64-bounded": "X64", "int128-bounded": "U128", "int128-checked": "C128"}[g.last_int_mode()], kernel_ms=kms, rows_kernels_ms=rms,
                        int32_ops_per_neighbor=w32, peak_source="tb_int_peak_detail microbenchmark, measured in this run: best of five "
                        "arms of single-SASS-instruction chains on all SMs, operations counted per instruction; "
                        "nominal = SMs x 4 schedulers x 32 lanes x SM clock (one warp instruction per scheduler per clock)",
                        peak_detail=peak, fp64_issue_arms=fp64_arms,
                        hbm=dict(achieved=hbm_achieved, peak=hbm_peak, unit="GB/s", frac=hbm_achieved / hbm_peak,
                                 peak_source="MEASURED_PEAKS.json" if peaks_file.exists() else "fallback"))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            primes = REF_SAMPLE_PRIMES[:min(cores, len(REF_SAMPLE_PRIMES))]
            n, s, kind = reference_sample(primes, precise=args.precise)
            cpu = dict(value=n / s, unit=UNIT, cores=len(primes), kind=kind,
                       sample=f"same genus, Genus::hecke_matrix_sparse(p) {'GMP' if args.precise else 'int64'} for the "
                              f"{len(primes)} primes {primes[0]}..{primes[-1]}, one process per prime; {s:.1f} s wall"
                              + fullsize_note(len(primes), n / s))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=g.last_int_mode() if args.precise else "int64", data="synthetic",
                    config=dict(workload=workload_name(p), neighbors_per_step_per_gpu=neighbors_per_step,
                                l2="flushed between timed steps (256 MB write); the 412 MB packed-word stream "
                                   "exceeds L2, the 3 MB genus table is L2-resident by design",
                                sharding="one T_p per rank per step (prime axis), no data-path collective"),
                    roofline=roofline, cpu_baseline=cpu,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=d2h_bytes,
                             result_bytes_per_step=result_bytes,
                             warmup_steps=e2e_warmup, step_wall_ms=my_steps,
                             slowest_rank_step_wall_ms=max((r["steps"] for r in per_rank), key=lambda v: sum(v[:-1])),
                             readback="per rank, chosen from measured delivery rates (tb_readback_stats): uint16 column + int8 "
                                      "entry over PCIe widened to the reference's int32 CSR by host threads, or int32 CSR over PCIe",
                             readback_policy=[r["policy"] for r in per_rank],
                             host_ingest=host_ingest_summary(per_rank, result_bytes / neighbors_per_step, e2e_value)),
                    gpu_launches=int(launches_timed), clocks=clocks, numa_node=numa_node, T_p_wall_ms=total_ms / args.steps,
                    final_gather_ms=(gather or {}).get("ms"), final_gather=gather)
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
