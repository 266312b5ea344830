"""CPU: host logic of the multi-GPU path -- LPT sharding, header rebasing, and the variable-size gather over
a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smoothxg_b200 import shard
from smoothxg_b200.engine import HDR_WORDS, H_OFF_HI, H_OFF_LO
from smoothxg_b200.synth import make_batch


def test_lpt_is_deterministic_balanced_partition():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 1000, 500).astype(float)
    for world in (1, 2, 4, 8):
        parts = shard.lpt_shard(costs, world)
        allb = np.sort(np.concatenate(parts))
        assert np.array_equal(allb, np.arange(500))
        loads = [costs[p].sum() for p in parts]
        assert max(loads) - min(loads) <= costs.max()
        again = shard.lpt_shard(costs, world)
        assert all(np.array_equal(a, b) for a, b in zip(parts, again))


def test_block_costs():
    b = make_batch(n_blocks=3, n_seqs=4, length=100, seed=1)
    c = shard.block_costs(b)
    tot = [int(b.block(i)[0].sum()) for i in range(3)]
    assert np.allclose(c, np.array(tot, dtype=float) ** 2)


def _fake_part(ids, words_each):
    h = np.zeros((len(ids), HDR_WORDS), dtype=np.int32)
    off = 0
    arena = []
    for k, b in enumerate(ids):
        h[k, 1] = 100 + b
        h[k, H_OFF_LO] = off
        arena.append(np.full(words_each, b, dtype=np.int32))
        off += words_each
    return np.asarray(ids), h.reshape(-1), np.concatenate(arena) if arena else np.zeros(0, np.int32)


def test_merge_parts_rebases_offsets():
    p0, p1 = _fake_part([0, 3], 5), _fake_part([1, 2, 4], 7)
    hdr, arena = shard.merge_parts(5, [p0, p1])
    H = hdr.reshape(-1, HDR_WORDS)
    for b in range(5):
        off = (int(H[b, H_OFF_LO]) & 0xFFFFFFFF) | (int(H[b, H_OFF_HI]) << 32)
        assert H[b, 1] == 100 + b and arena[off] == b


def test_rebase_local_multiple_arenas():
    ids, h, _ = _fake_part([0, 1, 2], 4)
    H = h.reshape(-1, HDR_WORDS).copy()
    H[2, H_OFF_LO] = 0  # block 2 was re-run into a second arena at offset 0
    out = shard.rebase_local(H.reshape(-1), [8, 4], np.array([0, 0, 1], dtype=np.int32)).reshape(-1, HDR_WORDS)
    assert out[0, H_OFF_LO] == 0 and out[1, H_OFF_LO] == 4 and out[2, H_OFF_LO] == 8


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = torch.arange(3 + 4 * rank, dtype=torch.int32) + 100 * rank
    got = shard.gather_to_root(t, dist)
    if rank == 0:
        q.put([g.tolist() for g in got])
    dist.barrier()
    dist.destroy_process_group()


def test_gather_to_root_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [list(range(3)), [100 + i for i in range(7)]]


@pytest.mark.gpu
def test_run_shard_world1_nccl_matches_run_batch():
    """The multi-GPU call path end to end on one GPU (world_size 1 NCCL group): H2D of the shard, kernel, the size-exact gather,
    poa_b200_result_from_device_parts (one D2H into pinned memory) == poa_b200_run_batch, block by block."""
    from smoothxg_b200 import engine as E
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(29700 + os.getpid() % 200)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        batch = make_batch(n_blocks=300, n_seqs=8, length=300, seed=9, indel_prob=0.2, indel_len=(5, 60))
        eng = E.PoaEngine(device=0)
        p = E.make_params(out_msa=True)
        ids = shard.plan(batch, 1)
        tm = {}
        got = shard.run_shard(eng, batch.select(ids[0]), ids, batch.n_blocks, p, dist, None, tm)
        want = eng.run_batch(batch, p)
        assert all(got.block_hash(b) == want.block_hash(b) for b in range(batch.n_blocks))
        assert np.array_equal(got.block(7).path_node, want.block(7).path_node) and np.array_equal(got.block(7).msa, want.block(7).msa)
        assert tm["gathered_bytes"] > 0 and "gather_ms" in tm
        got.close(); want.close(); eng.close()
    finally:
        dist.destroy_process_group()
