"""Multi-GPU driver for the POA path: static sharding of blocks over ranks and one gather at the end.

Blocks are independent POA problems (the reference's loop over them is a plain OpenMP parallel-for,
src/smooth.cpp:1931), so there is no exchange during compute.  Each rank (one process per GPU) aligns
its shard; the per-block result bodies then travel to rank 0 in ONE variable-size gather
(all_gather of sizes + gather of padded buffers; NCCL over NVLink on GPUs, gloo in the CPU tests) and are
re-ordered by block id.  The assignment is fixed before launch (cost-balanced LPT on sum(len)^2), so
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


def gather_to_root(t, dist, root: int = 0):
    """Variable-length gather of a 1-D tensor to `root`: one all_gather of sizes, one gather of padded buffers."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty(m, dtype=t.dtype, device=t.device) for _ in range(world)] if rank == root else None
    dist.gather(pad, bufs, dst=root)
    if rank != root:
        return None
    return [b[:s] for b, s in zip(bufs, sizes)]


class _DevPtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def run_sharded(eng, batch, params, dist=None, stream=None):
    """Align `batch` over all ranks of `dist` (None = single process).  Every rank passes the same batch;
    rank r aligns shard r.  Returns a PoaResult for the whole batch on rank 0 and None elsewhere."""
    import torch
    from . import engine
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    ids = lpt_shard(block_costs(batch), world)[rank]
    sub = batch.select(ids)
    dev = eng.upload(sub, params)
    dev.launch(stream); dev.finish(stream)
    d_hdr, d_arenas, block_arena = dev.device_result(sub.n_blocks)
    dv = torch.device("cuda", torch.cuda.current_device())
    hdr_t = torch.as_tensor(_DevPtr(d_hdr, max(sub.n_blocks, 1) * HDR_WORDS), device=dv)[:sub.n_blocks * HDR_WORDS]
    ar_t = [torch.as_tensor(_DevPtr(p, max(w, 1)), device=dv)[:w] for p, w in d_arenas]
    # offset fix-up is header-only and tiny, done on the host; bodies stay on the device
    h = rebase_local(hdr_t.cpu().numpy(), [w for _, w in d_arenas], block_arena)
    arena = torch.cat(ar_t) if len(ar_t) > 1 else (ar_t[0] if ar_t else torch.zeros(0, dtype=torch.int32, device=dv))
    if world == 1:
        a = arena.cpu().numpy()
        dev.close()
        hh, aa = merge_parts(batch.n_blocks, [(ids, h, a)])
        return engine.result_from_parts(hh, aa)
    payload = torch.cat([torch.from_numpy(np.concatenate([[ids.shape[0]], ids]).astype(np.int64)).view(torch.int32).to(dv),
                         torch.from_numpy(h).to(dv), arena])
    got = gather_to_root(payload, dist)
    dev.close()
    if rank != 0:
        return None
    parts = []
    for g in got:
        g = g.cpu().numpy()
        n_local = int(g[:2].view(np.int64)[0])
        gid = g[2:2 + 2 * n_local].view(np.int64)
        hh = g[2 + 2 * n_local:2 + 2 * n_local + n_local * HDR_WORDS]
        parts.append((gid, hh, g[2 + 2 * n_local + n_local * HDR_WORDS:]))
    hh, aa = merge_parts(batch.n_blocks, parts)
    return engine.result_from_parts(hh, aa)
