"""Where the end-to-end time of one poa_b200_run_batch goes (host wall clock per stage, 10 000-block shard)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import gen_batch
from smoothxg_b200 import engine
batch = gen_batch("10000x32x2kb", seed=1000)
import copy
pinned = copy.copy(batch); keep = []
for name in ("block_seq_off", "seq_len", "seq_off", "bases", "weight"):
    t = torch.from_numpy(getattr(batch, name)).pin_memory(); keep.append(t); setattr(pinned, name, t.numpy())
eng = engine.PoaEngine(device=0)
p = engine.make_params()
for it in range(3):
    t0 = time.perf_counter(); dev = eng.upload(pinned, p)
    t1 = time.perf_counter(); dev.launch()
    t2 = time.perf_counter(); dev.finish()
    t3 = time.perf_counter(); res = dev.download()
    t4 = time.perf_counter(); st = res.stats(); dev.close()
    t5 = time.perf_counter(); n = res.block(0).n_node; res.close()
    t6 = time.perf_counter()
    print(f"it{it}: upload {1e3*(t1-t0):.1f} launch {1e3*(t2-t1):.1f} finish {1e3*(t3-t2):.1f} download {1e3*(t4-t3):.1f} free {1e3*(t5-t4):.1f} view+free {1e3*(t6-t5):.1f} total {1e3*(t6-t0):.1f} ms | kernel {st['kernel_ms']:.1f} h2d {st['h2d_ms']:.1f} d2h {st['d2h_ms']:.1f} ms, d2h {st['d2h_bytes']/1e9:.2f} GB", flush=True)
for it in range(2):
    t0 = time.perf_counter(); r = eng.run_batch(pinned, p); t1 = time.perf_counter(); r.close()
    print(f"run_batch {1e3*(t1-t0):.1f} ms")
