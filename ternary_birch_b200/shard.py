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


_P2P_WARM = set()


def gather_hecke_device(result, device, group=None, dst: int = 0, widen: bool = True):
    """Final integer gather of device-resident Hecke rows (a `HeckeResult`) to rank `dst` over NCCL.

    Every rank packs all its conductors into ONE device buffer in the 3-byte form
    (`tb_hecke_pack`: uint16 column + int8 entry, int32 pairs when they do not fit), the
    sizes travel in one small all_gather, and the buffers go rank -> dst as grouped
    send/recv (`batch_isend_irecv`: one NCCL group, nothing padded, one receive buffer of the
    exact total).  On dst each rank's buffer is widened on the device into the reference's
    int32 CSR (`tb_unpack_rows`).  The first call per process pair also pays NCCL's
    point-to-point connection set-up; `warm_gather` does that outside a timed region.

    Returns on dst a list over ranks of {conductor: (data, indices, indptr)} int32 device
    tensors (views into one buffer per rank), plus the byte counts; None elsewhere.
    """
    import ctypes

    import torch
    import torch.distributed as dist

    from . import cabi

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    stream = torch.cuda.current_stream(device)
    C = result.num_conductors()
    nbytes, fmt = result.packed_bytes()
    meta = [nbytes, fmt]
    for k in range(C):
        rows, _, nnz = result.shape(k)
        meta += [result.conductor(k), nnz, rows]
    meta_t = torch.tensor(meta, dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta_t) for _ in range(world)]
    dist.all_gather(metas, meta_t, group=group)
    metas = torch.stack(metas).cpu().numpy()
    sizes = [int(m[0]) for m in metas]

    packed = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
    result.pack_to(packed.data_ptr(), stream.cuda_stream)
    ops, recv, offs = [], None, None
    if rank == dst:
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        recv = torch.empty(max(int(offs[-1]), 16), dtype=torch.uint8, device=device)
        recv[int(offs[dst]):int(offs[dst]) + nbytes].copy_(packed[:nbytes])
        for r in range(world):
            if r != dst and sizes[r] > 0:
                ops.append(dist.P2POp(dist.irecv, recv[int(offs[r]):int(offs[r + 1])], r, group))
    elif nbytes > 0:
        ops.append(dist.P2POp(dist.isend, packed[:nbytes], dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    out = []
    L = cabi.lib()
    for r in range(world):
        m = metas[r]
        conds = [int(m[2 + 3 * k]) for k in range(C)]
        nnz = np.ascontiguousarray([int(m[3 + 3 * k]) for k in range(C)], dtype=np.int64)
        rows = np.ascontiguousarray([int(m[4 + 3 * k]) for k in range(C)], dtype=np.int64)
        if not widen:
            out.append(dict(conductors=conds, nnz=nnz, rows=rows, entry_bytes=int(m[1]),
                            packed=recv[int(offs[r]):int(offs[r + 1])]))
            continue
        tot, totp = int(nnz.sum()), int((rows + 1).sum())
        buf = torch.empty(2 * tot + totp, dtype=torch.int32, device=device)
        base = buf.data_ptr()
        cabi.check(L.tb_unpack_rows(ctypes.c_void_p(recv.data_ptr() + int(offs[r])), int(m[1]), C, nnz.ctypes.data,
                                    rows.ctypes.data, ctypes.c_void_p(base), ctypes.c_void_p(base + 4 * tot),
                                    ctypes.c_void_p(base + 8 * tot), ctypes.c_void_p(stream.cuda_stream)))
        part, e, p_ = {}, 0, 0
        for k in range(C):
            n, np1 = int(nnz[k]), int(rows[k]) + 1
            part[conds[k]] = (buf[e:e + n], buf[tot + e:tot + e + n], buf[2 * tot + p_:2 * tot + p_ + np1])
            e += n
            p_ += np1
        out.append(part)
    return out, dict(bytes_per_rank=sizes, entry_bytes=[int(m[1]) for m in metas])


def warm_gather(device, group=None, dst: int = 0):
    """Set up NCCL's point-to-point connections rank -> dst once (they are made lazily on first
    use and cost far more than the gather itself), so that a later gather_hecke_device runs at
    link speed."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    key = (id(group), dst)
    if key in _P2P_WARM:
        return
    tiny = torch.zeros(4, dtype=torch.uint8, device=device)
    ops = []
    if rank == dst:
        bufs = [torch.zeros_like(tiny) for _ in range(world)]
        ops = [dist.P2POp(dist.irecv, bufs[r], r, group) for r in range(world) if r != dst]
    else:
        ops = [dist.P2POp(dist.isend, tiny, dst, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    torch.cuda.synchronize(device)
    _P2P_WARM.add(key)
