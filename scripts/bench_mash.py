#!/usr/bin/env python
"""Throughput of the identity estimate (include/mash_b200.h) on BASELINE configs[2]-shaped blocks, next to the reference's
own CPU implementation (unmodified mkmh/rkmh headers, oracle/_ref/libmash_ref.so, one thread per block over all host
threads) on a bounded sample.  Prints one JSON line.  usage: python scripts/bench_mash.py [--blocks N] [--steps K]"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from smoothxg_b200 import adaptive, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=2000)
ap.add_argument("--seqs", type=int, default=32)
ap.add_argument("--len", type=int, default=2000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--cpu-sample", type=int, default=64)
args = ap.parse_args()

batch = synth.make_batch(n_blocks=args.blocks, n_seqs=args.seqs, length=args.len, divergence=0.02, seed=1000)
fb = adaptive.from_codes(batch)
for _ in range(args.warmup):
    adaptive.block_identity(fb)
wall, stats = [], None
for _ in range(args.steps):
    t0 = time.perf_counter()
    r = adaptive.block_identity(fb)
    wall.append(time.perf_counter() - t0)
    stats = r["stats"]
t = sum(wall) / len(wall)
dev_ms = stats["hash_ms"] + stats["sort_ms"] + stats["compare_ms"]
line = {"metric": "adaptive_poa_identity_blocks_per_s", "value": args.blocks / t, "unit": "blocks/s", "ms_per_step": t * 1e3,
        "config": {"workload": f"synthetic {args.blocks} blocks x {args.seqs} seqs x {args.len} bp, 2% divergence, k=17 (BASELINE.json configs[2] shape)"},
        "device_ms": {k: round(stats[k], 3) for k in ("h2d_ms", "hash_ms", "sort_ms", "compare_ms", "d2h_ms", "host_ms")},
        "kernels_only_blocks_per_s": args.blocks / (dev_ms / 1e3), "n_hashes": stats["n_hashes"], "n_pairs": stats["n_pairs"],
        # algorithmic bytes: hash writes 8 B per k-mer; sort reads and writes each list once (16 B per hash); compare reads
        # every list (S-1) times from L2 (not HBM)
        "hbm_gbs": {"hash": stats["n_hashes"] * 9 / (stats["hash_ms"] / 1e3) / 1e9, "sort": stats["n_hashes"] * 16 / (stats["sort_ms"] / 1e3) / 1e9},
        "gpu_launches": stats["kernel_launches"]}
try:
    from oracle.mash import MashRef, ref_available
    if ref_available():
        ref = MashRef()
        n = min(args.cpu_sample, args.blocks)
        blocks = [fb.strings(b) for b in range(n)]
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL: one block per task, as the OpenMP loop does
            thr = list(ex.map(lambda s: ref.block(s, 17)[1], blocks))
        secs = time.perf_counter() - t0
        assert all(np.float32(a) == b for a, b in zip(thr, r["threshold"][:n])), "GPU thresholds differ from the reference"
        line["cpu_baseline"] = {"value": n / secs, "unit": "blocks/s", "cores": threads, "kind": "reference",
                                "sample": f"first {n} blocks, unmodified mkmh/rkmh (oracle/_ref/libmash_ref.so), one block per task, {secs:.1f} s; thresholds equal the GPU's"}
except Exception as e:  # the baseline is optional
    line["cpu_baseline"] = {"error": str(e)}
print(json.dumps(line))
