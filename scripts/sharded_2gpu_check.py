"""Multi-GPU check (run under torchrun, one rank per GPU): the statically sharded batch + one NCCL gather gives,
on rank 0, exactly what a single GPU gives for the whole batch (and what the oracle gives on a sample)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from smoothxg_b200 import engine, shard, synth
from tests.helpers import view_to_dump
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
batch = synth.make_batch(n_blocks=800, n_seqs=8, length=400, seed=77, indel_prob=0.2)
eng = engine.PoaEngine(device=lr)
p = engine.make_params(out_msa=True)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
res = shard.run_sharded(eng, batch, p, dist=dist)
torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
if rank == 0:
    one = eng.run_batch(batch, p)
    bad = sum(not np.array_equal(view_to_dump(res.block(b)).result_part(), view_to_dump(one.block(b)).result_part()) for b in range(batch.n_blocks))
    from oracle.oracle import Oracle, make_params as op
    ora = Oracle()
    bad_o = sum(not np.array_equal(view_to_dump(res.block(b)).result_part(), ora.poa_block(op(out_msa=True), *batch.block(b), instrument=False).result_part()) for b in range(0, batch.n_blocks, 97))
    print(f"world={world}: sharded run {1e3*(t1-t0):.0f} ms; blocks differing from the single-GPU run: {bad}; from the oracle (sample): {bad_o}", flush=True)
    assert bad == 0 and bad_o == 0
dist.destroy_process_group()
