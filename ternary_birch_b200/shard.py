"""Multi-GPU partitioning of the p-neighbor path.

Row n of every T_p^(k) depends only on representative n's p+1 neighbors and the
read-only genus table (Genus.h:562-660), and different primes are independent
calls, so the work shards as (representative range x prime) with NO collective
on the hot path.  The only exchange is a final integer gather of the Hecke rows
to rank 0, done here with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def row_shards(num_reps: int, world: int):
    """Contiguous, balanced blocks of representatives: [(begin, end)] * world."""
    base, extra = divmod(num_reps, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def prime_shards(primes, world: int):
    """Deal primes round-robin by descending p (cost per T_p grows with p)."""
    order = sorted(primes, reverse=True)
    return [order[r::world] for r in range(world)]


def concat_row_shards(parts):
    """parts: per-rank {conductor: (data, indices, indptr)} in rank order -> the full CSR per conductor.

    Row blocks are contiguous, so data/indices concatenate and indptr is offset by the
    running nnz (the reference appends rows in order, Genus.h:643-656).
    """
    out = {}
    for cond in parts[0]:
        data = np.concatenate([np.asarray(p[cond][0], dtype=np.int32) for p in parts])
        indices = np.concatenate([np.asarray(p[cond][1], dtype=np.int32) for p in parts])
        ptr = [np.zeros(1, dtype=np.int64)]
        off = 0
        for p in parts:
            ip = np.asarray(p[cond][2], dtype=np.int64)
            ptr.append(ip[1:] + off)
            off += int(ip[-1])
        out[cond] = (data, indices, np.concatenate(ptr).astype(np.int32))
    return out


def pack_csr(local: dict):
    """{conductor: (data, indices, indptr)} -> (header int64[3*C+1], payload int32) for one transfer."""
    conds = sorted(local)
    header = [len(conds)]
    chunks = []
    for c in conds:
        data, indices, indptr = local[c]
        header += [int(c), int(len(data)), int(len(indptr))]
        chunks += [np.asarray(data, dtype=np.int32), np.asarray(indices, dtype=np.int32),
                   np.asarray(indptr, dtype=np.int32)]
    payload = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int32)
    return np.asarray(header, dtype=np.int64), payload


def unpack_csr(header, payload):
    header = [int(x) for x in header]
    out, pos, h = {}, 0, 1
    for _ in range(header[0]):
        c, nnz, nptr = header[h:h + 3]
        h += 3
        data = payload[pos:pos + nnz]; pos += nnz
        indices = payload[pos:pos + nnz]; pos += nnz
        indptr = payload[pos:pos + nptr]; pos += nptr
        out[c] = (np.asarray(data), np.asarray(indices), np.asarray(indptr))
    return out


def gather_csr(local: dict, device=None, group=None, dst: int = 0):
    """Final integer gather of every rank's Hecke rows to `dst` (returns the list in rank order
    on dst, None elsewhere).  int32 payloads travel as one padded tensor per rank."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    header, payload = pack_csr(local)
    dev = device if device is not None else torch.device("cpu")
    sizes = torch.tensor([header.size, payload.size], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    max_h = max(int(s[0]) for s in all_sizes)
    max_p = max(int(s[1]) for s in all_sizes)
    hbuf = torch.zeros(max_h, dtype=torch.int64, device=dev)
    hbuf[:header.size] = torch.from_numpy(header).to(dev)
    pbuf = torch.zeros(max(max_p, 1), dtype=torch.int32, device=dev)
    pbuf[:payload.size] = torch.from_numpy(payload).to(dev)
    hs = [torch.zeros_like(hbuf) for _ in range(world)] if rank == dst else None
    ps = [torch.zeros_like(pbuf) for _ in range(world)] if rank == dst else None
    dist.gather(hbuf, hs, dst=dst, group=group)
    dist.gather(pbuf, ps, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        nh, npay = int(all_sizes[r][0]), int(all_sizes[r][1])
        out.append(unpack_csr(hs[r][:nh].cpu().numpy(), ps[r][:npay].cpu().numpy()))
    return out


def gather_hecke_device(result, device, group=None, dst: int = 0):
    """Final integer gather of device-resident Hecke rows (a `HeckeResult`) to rank `dst`.

    Each rank lays its CSR arrays out in ONE int32 device tensor (copied device-to-device out
    of the library's buffers), the tensors are padded to the largest rank and gathered over
    NCCL/NVLink.  Returns on dst: (headers, payloads) with headers[r] = int64 [C, then per
    conductor: conductor, nnz, rows+1] and payloads[r] an int32 device tensor holding, per
    conductor, data | indices | indptr; None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    C = result.num_conductors()
    header = [C]
    total = 0
    for k in range(C):
        rows, _, nnz = result.shape(k)
        header += [result.conductor(k), nnz, rows + 1]
        total += 2 * nnz + rows + 1
    payload = torch.empty(max(total, 1), dtype=torch.int32, device=device)
    pos = 0
    base = payload.data_ptr()
    for k in range(C):
        rows, _, nnz = result.shape(k)
        result.copy_sparse_to(k, base + 4 * pos, base + 4 * (pos + nnz), base + 4 * (pos + 2 * nnz))
        pos += 2 * nnz + rows + 1
    hdr = torch.tensor(header + [total], dtype=torch.int64, device=device)
    hdrs = [torch.zeros_like(hdr) for _ in range(world)]
    dist.all_gather(hdrs, hdr, group=group)
    max_total = max(int(h[-1]) for h in hdrs)
    if payload.numel() < max_total:
        padded = torch.zeros(max_total, dtype=torch.int32, device=device)
        padded[:payload.numel()] = payload
        payload = padded
    else:
        payload = payload[:max_total] if max_total > 0 else payload[:1]
    bufs = [torch.empty_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(payload, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [h[:-1].cpu().numpy() for h in hdrs], [b[:int(h[-1])] for b, h in zip(bufs, hdrs)]
