"""Multi-GPU driver for the POA path: static sharding of blocks over ranks and one gather at the end.

Blocks are independent POA problems (the reference's loop over them is a plain OpenMP parallel-for,
src/smooth.cpp:1931), so there is no exchange during compute.  Each rank (one process per GPU) aligns
its shard; the per-block result bodies then travel to rank 0 in ONE variable-size gather
(all_gather of sizes + one grouped, size-exact send/recv into a flat buffer; NCCL over NVLink on GPUs, gloo in the
CPU tests) and are re-ordered by block id through their headers.  The assignment is fixed before launch (cost-balanced LPT on sum(len)^2), so
results are bit-identical for any GPU count.
"""
from __future__ import annotations

import heapq

import numpy as np

from .engine import HDR_WORDS, H_OFF_HI, H_OFF_LO


def block_costs(batch) -> np.ndarray:
    """Cost model of one block: DP cells grow with (total bases)^2 / n_seq * ... ~ (sum of lengths)^2."""
    tot = np.diff(batch.seq_off[batch.block_seq_off].astype(np.int64)).astype(np.float64)
    return tot * tot


def lpt_shard(costs, world: int) -> list:
    """Longest-processing-time-first greedy assignment; returns world arrays of block ids (ascending)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(costs.shape[0]), -costs))  # cost desc, id asc: deterministic
    heap = [(0.0, r) for r in range(world)]
    heapq.heapify(heap)
    out = [[] for _ in range(world)]
    for b in order:
        load, r = heapq.heappop(heap)
        out[r].append(int(b))
        heapq.heappush(heap, (load + float(costs[b]), r))
    return [np.array(sorted(x), dtype=np.int64) for x in out]


def merge_parts(n_blocks_total: int, parts) -> tuple:
    """parts: iterable of (block_ids, hdr[int32 n_local*HDR_WORDS], arena[int32]) per rank -> (hdr, arena) of the
    whole batch with body offsets rebased onto the concatenated arena."""
    hdr = np.full(n_blocks_total * HDR_WORDS, -1, dtype=np.int32)
    arenas, base = [], 0
    for ids, h, a in parts:
        h = np.asarray(h, dtype=np.int32).reshape(-1, HDR_WORDS).copy()
        off = (h[:, H_OFF_LO].astype(np.int64) & 0xFFFFFFFF) | (h[:, H_OFF_HI].astype(np.int64) << 32)
        off += base
        h[:, H_OFF_LO] = (off & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
        h[:, H_OFF_HI] = (off >> 32).astype(np.int32)
        hdr.reshape(-1, HDR_WORDS)[np.asarray(ids, dtype=np.int64)] = h
        arenas.append(np.asarray(a, dtype=np.int32))
        base += arenas[-1].shape[0]
    return hdr, (np.concatenate(arenas) if arenas else np.zeros(0, np.int32))


def rebase_local(hdr: np.ndarray, arena_words: list, block_arena: np.ndarray) -> np.ndarray:
    """A rank's result may sit in several arenas (re-run blocks).  Returns the headers with body offsets
    rebased onto the concatenation of those arenas."""
    if len(arena_words) <= 1:
        return hdr
    h = hdr.reshape(-1, HDR_WORDS).copy()
    bases = np.concatenate([[0], np.cumsum(arena_words)[:-1]]).astype(np.int64)
    off = (h[:, H_OFF_LO].astype(np.int64) & 0xFFFFFFFF) | (h[:, H_OFF_HI].astype(np.int64) << 32)
    ok = block_arena >= 0
    off[ok] += bases[block_arena[ok]]
    h[:, H_OFF_LO] = (off & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    h[:, H_OFF_HI] = (off >> 32).astype(np.int32)
    return h.reshape(-1)


def gather_exact(t, dist, root: int = 0):
    """Size-exact variable-length gather of a 1-D tensor to `root`: one all_gather of element counts, then ONE grouped
    exchange -- every other rank sends exactly its payload, `root` receives each straight into its slice of one flat
    buffer (nothing is padded to the largest payload, nothing is copied a second time on the device).  Returns
    (flat buffer, [start offsets], [sizes]) on root, None elsewhere.  NCCL over NVLink on GPUs, gloo in the CPU tests."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    starts = [0] * world
    for r in range(1, world):
        starts[r] = starts[r - 1] + sizes[r - 1]
    if rank != root:
        if t.numel():
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, t, root)]):
                w.wait()
        return None
    buf = torch.empty(max(sum(sizes), 1), dtype=t.dtype, device=t.device)
    ops = [dist.P2POp(dist.irecv, buf[starts[r]:starts[r] + sizes[r]], r) for r in range(world) if r != root and sizes[r]]
    works = dist.batch_isend_irecv(ops) if ops else []
    buf[starts[root]:starts[root] + sizes[root]].copy_(t)
    for w in works:
        w.wait()
    return buf, starts, sizes


def gather_to_root(t, dist, root: int = 0):
    """Variable-length gather of a 1-D tensor to `root` as a list of per-rank tensors (views of gather_exact()'s buffer)."""
    got = gather_exact(t, dist, root)
    if got is None:
        return None
    buf, starts, sizes = got
    return [buf[a:a + n] for a, n in zip(starts, sizes)]


class _DevPtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def plan(batch, world: int) -> list:
    """The static assignment: block ids of every rank (ascending), fixed before launch."""
    return lpt_shard(block_costs(batch), world)


def gather_result(eng, dev, n_local: int, ids_by_rank, n_total: int, dist, stream=None, timings=None):
    """The path's one exchange step: a finished device batch of this rank's shard -> PoaResult of the whole batch on rank 0
    (None elsewhere).  Payload per rank = [headers | result bodies], exactly as the kernel left them in HBM; rank 0
    receives every payload into one flat device buffer over NCCL, rebases the (tiny) headers on the host, and builds the
    host result with a single device-to-host copy into pooled pinned memory (poa_b200_result_from_device_parts)."""
    import time
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    d_hdr, d_arenas, block_arena = dev.device_result(n_local)
    dv = torch.device("cuda", torch.cuda.current_device())
    hdr_t = torch.as_tensor(_DevPtr(d_hdr, max(n_local, 1) * HDR_WORDS), device=dv)[:n_local * HDR_WORDS]
    ar_t = [torch.as_tensor(_DevPtr(p, max(w, 1)), device=dv)[:w] for p, w in d_arenas]
    if len(ar_t) > 1:  # re-run blocks sit in further arenas: rebase their offsets onto the concatenation (headers only, host)
        h = rebase_local(hdr_t.cpu().numpy(), [w for _, w in d_arenas], block_arena)
        hdr_t = torch.from_numpy(h).to(dv)
    payload = torch.cat([hdr_t] + ar_t) if (n_local or ar_t) else torch.zeros(0, dtype=torch.int32, device=dv)
    t0 = time.perf_counter()
    got = gather_exact(payload, dist)
    if timings is not None:
        torch.cuda.synchronize()
        timings["gather_ms"] = (time.perf_counter() - t0) * 1e3
    if rank != 0:
        return None
    buf, starts, sizes = got
    # headers of every rank: small strided device reads, rebased onto the flat buffer on the host
    hdr = np.full(n_total * HDR_WORDS, -1, dtype=np.int32)
    H = hdr.reshape(-1, HDR_WORDS)
    for r in range(world):
        ids = np.asarray(ids_by_rank[r], dtype=np.int64)
        nl = ids.shape[0]
        if nl == 0:
            continue
        h = buf[starts[r]:starts[r] + nl * HDR_WORDS].cpu().numpy().reshape(-1, HDR_WORDS).copy()
        off = (h[:, H_OFF_LO].astype(np.int64) & 0xFFFFFFFF) | (h[:, H_OFF_HI].astype(np.int64) << 32)
        off += starts[r] + nl * HDR_WORDS
        h[:, H_OFF_LO] = (off & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
        h[:, H_OFF_HI] = (off >> 32).astype(np.int32)
        H[ids] = h
    t1 = time.perf_counter()
    res = eng.result_from_device(hdr, buf.data_ptr(), sum(sizes), stream)
    if timings is not None:
        timings["d2h_ms"] = (time.perf_counter() - t1) * 1e3
        timings["gathered_bytes"] = 4 * sum(sizes)
    return res


def run_shard(eng, sub, ids_by_rank, n_total: int, params, dist, stream=None, timings=None):
    """One rank's part of a sharded call: host shard in (`sub` = batch.select(ids_by_rank[rank])), H2D, kernel, the gather;
    the whole batch's PoaResult out on rank 0."""
    import time
    t0 = time.perf_counter()
    dev = eng.upload(sub, params)
    dev.launch(stream); dev.finish(stream)
    if timings is not None:
        st = dev.stats()
        timings["h2d_ms"] = st["h2d_ms"]; timings["kernel_ms"] = st["kernel_ms"]; timings["compute_ms"] = (time.perf_counter() - t0) * 1e3
        timings["h2d_bytes"] = st["h2d_bytes"]
    res = gather_result(eng, dev, sub.n_blocks, ids_by_rank, n_total, dist, stream, timings)
    dev.close()
    return res


def run_sharded(eng, batch, params, dist=None, stream=None):
    """Align `batch` over all ranks of `dist` (None = single process).  Every rank passes the same batch;
    rank r aligns shard r.  Returns a PoaResult for the whole batch on rank 0 and None elsewhere."""
    if dist is None or dist.get_world_size() == 1:
        return eng.run_batch(batch, params)
    ids_by_rank = plan(batch, dist.get_world_size())
    return run_shard(eng, batch.select(ids_by_rank[dist.get_rank()]), ids_by_rank, batch.n_blocks, params, dist, stream)
