#!/usr/bin/env python
"""bench.py -- POA hot-path benchmark (BASELINE.json metric: POA DP Gcells/s and blocks/s).

    python bench.py --gpus 1 --steps 3 --warmup 3            # our arm, one B200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                      # unmodified vendored abPOA on the host cores

Workload (config.workload): BASELINE.json configs[2], the configuration the north star quotes the
target on -- synthetic 10 000 blocks x 32 sequences x 2 kb, 2 % divergence, global alignment with
abPOA's adaptive band (wb=311, wf=0.03), convex gaps 1,4,6,2,26,1.  A "step" is one pass of the whole
per-block loop (DP fill, traceback, graph fusion, topological sort, consensus) over the batch.
STRONG scaling, as the north star asks: the ONE seed-1000 batch is sharded statically (cost-balanced LPT,
smoothxg_b200/shard.py) over the N ranks, and the per-block results travel to rank 0 in one NCCL gather.

  value  = in-band DP cells of the batch / max-over-ranks device time of the ranks' shards, inputs resident
           in HBM (cells counted by the kernel; identical to the oracle's band, see tests).
  e2e    = the same through the public call: N = 1 poa_b200_run_batch() (pinned host buffers in, host result
           out, H2D + kernels + D2H inside the timed region); N > 1 shard.run_shard() on every rank (H2D of
           the shard, kernels, the NCCL gather, ONE D2H on rank 0, the whole batch's host result on rank 0).
  weak   = (N > 1 only, extra key) every rank aligns the whole 10 000-block batch: the round-1 figure.
  parity_sample = the blocks of the CPU sample hashed (graph: node count, bases, out edges, weights) on both
           sides: unmodified abPOA vs this run's GPU result; a mismatch fails the run.
  roofline = algorithmic HBM bytes (sizeof(score) x (5 + 3 p-bar) per in-band cell) /
           kernel time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline = oracle/_ref (unmodified abPOA; AVX-512BW where the host has it, and AVX2 beside it) on all
           host threads over a bounded sample of the same batch (rank 0, N=1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n_blocks, n_seqs, length, divergence)
    "10000x32x2kb": (10000, 32, 2000, 0.02),
    "1000x16x1kb": (1000, 16, 1000, 0.02),
    "100x256x8kb": (100, 256, 8000, 0.02),   # BASELINE.json configs[3], deep-block stress (int32 scores once rows > 16 361)
    # SURVEY 8(d) variants of the headline batch (not headline lines): the -a preset regimes and real-block-like long indels
    "10000x32x2kb_d0.1": (10000, 32, 2000, 0.001),
    "10000x32x2kb_d5": (10000, 32, 2000, 0.05),
    "10000x32x2kb_indel": (10000, 32, 2000, 0.02),   # + one 50-500 bp insertion or deletion per copy with probability 0.3
    "8x64x4kb_probe": (8, 64, 4000, 0.02),           # profiling probe: deep-block behaviour (one block per SM, 8 warps) in a kernel short enough for ncu
}
EXTRA = {"10000x32x2kb_indel": dict(indel_prob=0.3, indel_len=(50, 500))}
METRIC = "poa_dp_inband_gcells_per_s"
UNIT = "Gcells/s"


def gen_batch(workload: str, seed: int, n_blocks: int | None = None):
    from smoothxg_b200 import synth
    nb, ns, L, d = WORKLOADS[workload]
    if n_blocks is not None:
        nb = n_blocks
    cache = os.path.join("/tmp", f"poa_bench_{workload}_{nb}_{seed}.npz")
    if os.path.exists(cache):
        z = np.load(cache)
        return synth.PoaBatch(z["bso"], z["sl"], z["so"], z["ba"], z["wt"])
    b = synth.make_batch(n_blocks=nb, n_seqs=ns, length=L, divergence=d, seed=seed, **EXTRA.get(workload, {}))
    try:
        tmp = f"{cache}.{os.getpid()}.tmp.npz"
        np.savez(tmp, bso=b.block_seq_off, sl=b.seq_len, so=b.seq_off, ba=b.bases, wt=b.weight)
        os.replace(tmp, cache)
    except OSError:
        pass
    return b


def host_cpu_info() -> dict:
    """CPU model, logical threads and physical cores of this host (the reference arm's denominator)."""
    model, pairs, phys, core = "unknown", set(), None, None
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                k, _, v = line.partition(":")
                k, v = k.strip(), v.strip()
                if k == "model name":
                    model = v
                elif k == "physical id":
                    phys = v
                elif k == "core id":
                    core = v
                elif not k and phys is not None and core is not None:
                    pairs.add((phys, core)); phys = core = None
    except OSError:
        pass
    threads = os.cpu_count() or 1
    try:
        threads = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    return {"model": model, "threads": threads, "physical_cores": len(pairs) or threads}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_per_launch(workload: str, cells_per_launch: float):
    """DRAM bytes per launch of the POA kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    `ncu --set full` capture of this workload.  The capture holds one launch over fewer blocks of the same shape
    (a full 10 000-block launch does not fit ncu's 40-pass replay in the GPU budget), so profiles/traffic.json
    stores bytes per in-band cell and the figure is scaled to this launch's cells."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        if t.get("workload") == workload and t.get("dram_bytes_per_cell"):
            return float(t["dram_bytes_per_cell"]) * cells_per_launch
    return None


def cpu_reference(batch, params_kw, n_sample: int, threads: int, isa: str | None = None, want_hash: bool = False):
    """Unmodified vendored abPOA (oracle/_ref) over the first n_sample blocks, OpenMP dynamic loop over
    blocks like reference src/smooth.cpp:1931.  Returns (sample, seconds, kind, simd, per-block graph hashes or None)."""
    from oracle.oracle import RefAbpoa, Oracle, make_params, ref_available
    sub = batch.select(range(min(n_sample, batch.n_blocks)))
    p = make_params(**params_kw)
    if ref_available(isa):
        ref = RefAbpoa(isa)
        ref.batch_timed(p, batch.select(range(min(threads, sub.n_blocks))), n_threads=threads)  # warm-up pass
        if want_hash:
            secs, h = ref.batch_timed(p, sub, n_threads=threads, want_hash=True)
            return sub, secs, "reference", ref.simd, h
        return sub, ref.batch_timed(p, sub, n_threads=threads), "reference", ref.simd, None
    if isa is not None:
        return sub, None, None, None, None
    ora = Oracle()  # scalar port, single thread
    t0 = time.perf_counter()
    ora.poa_batch(p, sub, instrument=False)
    return sub, time.perf_counter() - t0, "port", "scalar", None


def identity_estimate_leg(batch, device: int, n_blocks: int = 2000) -> dict:
    """The step in front of the POA (include/mash_b200.h): the --adaptive-poa-params identity estimate on
    the first n_blocks blocks of the same shard, host strings in, thresholds out; outside the timed POA region and never
    fatal for the headline line."""
    try:
        from smoothxg_b200 import adaptive
        fb = adaptive.from_codes(batch.select(range(min(n_blocks, batch.n_blocks))))
        adaptive.block_identity(fb, device=device)
        t0 = time.perf_counter()
        r = adaptive.block_identity(fb, device=device)
        secs = time.perf_counter() - t0
        stt = r["stats"]
        return {"value": fb.n_blocks / secs, "unit": "blocks/s", "blocks": fb.n_blocks, "ms": secs * 1e3,
                "device_ms": {k: round(stt[k], 2) for k in ("h2d_ms", "hash_ms", "sort_ms", "compare_ms", "d2h_ms", "host_ms")},
                "gpu_launches": stt["kernel_launches"], "min_threshold": float(r["threshold"].min()), "max_threshold": float(r["threshold"].max())}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def workload_text(workload: str, nb: int, ns: int, L: int, div: float) -> str:
    if workload == "10000x32x2kb":
        return (f"synthetic {nb} blocks x {ns} seqs x {L} bp, {div:.0%} divergence, global, adaptive band wb=311 wf=0.03, "
                f"convex gaps 1,4,6,2,26,1 (BASELINE.json configs[2])")
    if workload in ("1000x16x1kb", "100x256x8kb"):
        return f"synthetic {nb} blocks x {ns} seqs x {L} bp (BASELINE.json configs[{1 if workload == '1000x16x1kb' else 3}])"
    return (f"synthetic {nb} blocks x {ns} seqs x {L} bp, {div:.1%} divergence{', long indels' if workload in EXTRA else ''} "
            f"(SURVEY 8d variant of configs[2], not a headline line)")


def reference_arm(args, config, params_kw):
    """bench.py --impl reference: the unmodified abPOA on this box's host threads (rank 0 only)."""
    cpu = host_cpu_info()
    threads = cpu["threads"]
    batch = gen_batch(args.workload, seed=1000, n_blocks=args.blocks)
    n_sample = args.cpu_sample or max(threads, min(batch.n_blocks, 32 * threads))
    from oracle.oracle import Oracle, make_params
    times = []
    sub = kind = simd = None
    for it in range(args.warmup + args.steps):
        sub, secs, kind, simd, _ = cpu_reference(batch, params_kw, n_sample, threads)
        if it >= args.warmup:
            times.append(secs)
    # in-band cells of the sample from the oracle restatement (identical to abPOA's band, tests/test_oracle_vs_ref.py)
    ora = Oracle()
    n_cells = min(sub.n_blocks, 2 * threads)
    cells = sum(d.inband_cells for d in ora.poa_batch(make_params(**params_kw), sub.select(range(n_cells))))
    cells = cells * sub.n_blocks / n_cells
    t = sum(times) / len(times)
    val = cells / t / 1e9
    base = {"value": val, "unit": UNIT, "cores": threads, "physical_cores": cpu["physical_cores"], "cpu_model": cpu["model"], "kind": kind,
            "sample": f"first {sub.n_blocks} blocks of the batch per step, abPOA v1.5.4 {simd}, OpenMP dynamic over blocks, "
                      f"cells extrapolated from the oracle's band on {n_cells} blocks"}
    # the narrower ISA beside it (BASELINE metric: abPOA AVX2 and AVX-512); not the ratio's denominator
    s2, secs2, _, simd2, _ = cpu_reference(batch, params_kw, max(threads, n_sample // 2), threads, isa="avx2")
    if secs2:
        base["avx2"] = {"value": cells / sub.n_blocks * s2.n_blocks / secs2 / 1e9, "unit": UNIT, "blocks_per_s": s2.n_blocks / secs2, "simd": simd2, "blocks": s2.n_blocks}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic", "config": config, "blocks_per_s": sub.n_blocks / t, "cpu_baseline": base,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_RESULT_FD = None


def emit(line: dict) -> None:
    """The one JSON line of the contract, on the process's ORIGINAL stdout (see main(): libraries are kept off it)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="10000x32x2kb", choices=sorted(WORKLOADS))
    ap.add_argument("--blocks", type=int, default=None, help="override the batch's block count (debug; invalidates the headline)")
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="blocks in the CPU baseline sample (0 = auto)")
    ap.add_argument("--slab-rows-factor", type=float, default=0.0, help="DP workspace rows per query base (0 = engine default; tuning experiments)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaling figure")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the bit-identity check of the gathered result against a single-GPU run")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the result: NCCL prints its version banner on file descriptor 1 when the first communicator
    # is created, so everything else written to fd 1 from here on (libraries included) goes to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nb, ns, L, div = WORKLOADS[args.workload]
    if args.blocks:
        nb = args.blocks
    params_kw = dict(local=False, banded=True, out_cons=True, out_msa=False)
    config = {"workload": workload_text(args.workload, nb, ns, L, div),
              "blocks": nb, "seqs_per_block": ns, "seq_len": L, "divergence": div,
              "l2": "inputs + per-block workspaces are tens of GB per step, far larger than the 126 MB L2 (no flush needed)",
              "sharding": "ONE batch, static cost-balanced LPT shards over the ranks (fixed before launch), one NCCL gather of headers + result bodies to rank 0"}

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, config, params_kw)
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    from smoothxg_b200 import engine, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the POA engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        if rank == 0:
            gen_batch(args.workload, seed=1000, n_blocks=args.blocks)  # one rank writes the cache, the others read it
        dist.barrier()
    full = gen_batch(args.workload, seed=1000, n_blocks=args.blocks)
    ids_by_rank = shard.plan(full, world)
    batch = full.select(ids_by_rank[rank]) if world > 1 else full
    eng = engine.PoaEngine(device=local_rank, warps_per_block=args.warps, ctas_per_sm=args.ctas_per_sm, slab_rows_factor=args.slab_rows_factor)
    params = engine.make_params(**params_kw)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_leg(b):
        """K timed launch+finish passes over device-resident inputs; returns (ms, kernel_ms, launches, stats)."""
        dev = eng.upload(b, params)
        for _ in range(args.warmup):
            dev.launch(stream); dev.finish(stream)
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ms, launches = 0.0, 0
        e0.record()
        for _ in range(args.steps):
            dev.launch(stream); dev.finish(stream)
            st = dev.stats()
            kernel_ms += st["kernel_ms"]; launches += st["kernel_launches"]
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        st = dev.stats()
        dev.close()
        return ms, kernel_ms, launches, st, clocks

    def reduce_max(*vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(*vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.tolist()

    # --- device-resident leg over this rank's shard (N = 1: the whole batch)
    ms, kernel_ms, launches, st, clocks = device_leg(batch)
    cells, edge_rows = st["inband_cells"], st["edge_row_cells"]

    # --- end-to-end leg: host buffers in, the whole batch's host result out (on rank 0)
    def pin(b):
        import copy
        pb = copy.copy(b)
        keep = []
        for name in ("block_seq_off", "seq_len", "seq_off", "bases", "weight"):
            t = torch.from_numpy(getattr(b, name)).pin_memory()
            keep.append(t)
            setattr(pb, name, t.numpy())
        pb._keep = keep
        return pb

    e2e_ms = None
    last = None
    h2d = d2h = 0
    e2e_each, e2e_parts = [], {}
    verify = None
    if not args.no_e2e:
        pinned = pin(batch)

        def e2e_call(tm):
            if world == 1:
                r = eng.run_batch(pinned, params)
                s2 = r.stats()
                tm.update(h2d_ms=s2["h2d_ms"], kernel_ms=s2["kernel_ms"], d2h_ms=s2["d2h_ms"], h2d_bytes=s2["h2d_bytes"], d2h_bytes=s2["d2h_bytes"])
                return r
            r = shard.run_shard(eng, pinned, ids_by_rank, full.n_blocks, params, dist, stream, tm)
            if r is not None:
                tm["d2h_bytes"] = r.stats()["d2h_bytes"]
            return r

        for _ in range(2):  # warm the pinned/device pools (the first call page-locks GBs for the result)
            r = e2e_call({})
            if r is not None:
                r.close()
        barrier()
        t0 = time.perf_counter()
        chk = 3
        for _ in range(args.steps):
            ts = time.perf_counter()
            if last is not None:
                last.close()
            tm = {}
            last = e2e_call(tm)
            if last is not None:
                chk = last.block(0).n_node  # the step's result is read on the host
                d2h = tm.get("d2h_bytes", 0)
            h2d = tm.get("h2d_bytes", 0)
            e2e_each.append((time.perf_counter() - ts) * 1e3)
            for k in ("h2d_ms", "kernel_ms", "gather_ms", "d2h_ms"):
                if k in tm:
                    e2e_parts.setdefault(k, []).append(round(tm[k], 1))
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        assert chk > 2
        # the gathered result must be what ONE GPU gives for the whole batch, bit for bit (graph hash of every block)
        if world > 1 and not args.no_verify:
            if rank == 0:
                one = eng.run_batch(full, params)
                bad = sum(one.block_hash(b) != last.block_hash(b) for b in range(full.n_blocks))
                verify = {"blocks": full.n_blocks, "differing_from_single_gpu_run": int(bad)}
                one.close()
            dist.barrier()

    # --- N > 1: the round-1 weak figure beside it (every rank aligns the whole batch)
    weak = None
    if world > 1 and not args.no_weak:
        wms, _, _, wst, _ = device_leg(full)
        (wms,) = reduce_max(wms)
        weak = {"value": float(wst["inband_cells"]) * world * args.steps / (wms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": wms / args.steps,
                "what": "every rank aligns the whole batch (per-GPU work fixed), no data-path collective"}

    # --- max over ranks, sums over ranks
    ms, e2e_max, kernel_ms_max = reduce_max(ms, e2e_ms or 0.0, kernel_ms)
    tot_cells, tot_blocks, edge_rows_all, launches_all, h2d_all = reduce_sum(float(cells), float(batch.n_blocks), float(edge_rows), float(launches), float(h2d))
    e2e_ms = e2e_max if e2e_ms is not None else None

    if rank == 0:
        K = args.steps
        sec = ms / 1e3
        value = tot_cells * K / sec / 1e9
        pbar = edge_rows_all / max(tot_cells, 1.0)
        bytes_per_cell = 2.0 * (5.0 + 3.0 * pbar)  # int16 convex: write 5 planes, read 3 per predecessor edge
        peak, peak_src = peaks()
        per_launch_s = (kernel_ms / 1e3) / max(K, 1)  # this rank's POA kernel time per step (main launch plus re-runs of overflowed blocks)
        achieved = (float(cells) * bytes_per_cell) / per_launch_s / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                # dtype: also for deep blocks -- 16-bit cells provably hold every real value there (p16_safe_for_long_graph); abPOA itself would use int32
                "dtype": "int16", "data": "synthetic", "config": config,
                "blocks_per_s": tot_blocks * K / sec, "inband_cells_per_step": tot_cells, "p_bar": pbar,
                "clocks": clocks, "gpu_launches": int(launches_all),
                "engine": {"n_ctas": st["n_ctas"], "warps_per_block": st["warps_per_block"], "workspace_gb": st["workspace_bytes"] / 1e9,
                           "retried_blocks": st["retried_blocks"], "phase_cycles": st["phase_cycles"], "blocks_rank0": batch.n_blocks},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic_per_launch(args.workload, float(cells)), "peak_source": peak_src,
                             "kernel": "poa_b200_block_kernel", "algorithmic_bytes_per_cell": bytes_per_cell,
                             "kernel_ms_per_launch": per_launch_s * 1e3, "kernel_ms_per_launch_max_rank": kernel_ms_max / max(K, 1)}}
        if e2e_ms is not None:
            line["e2e"] = {"value": tot_cells * K / (e2e_ms / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h),
                           "blocks_per_s": tot_blocks * K / (e2e_ms / 1e3), "ms_per_step": e2e_ms / K,
                           "ms_each_step_rank0": [round(x, 1) for x in e2e_each], "parts_each_step_rank0": e2e_parts,
                           "path": "poa_b200_run_batch (pinned host in, host result out)" if world == 1 else
                                   "shard.run_shard on every rank: H2D of the shard, kernel, NCCL gather of headers + bodies to rank 0, one D2H, host result of the whole batch"}
        if verify is not None:
            line["gather_verify"] = verify
        if weak is not None:
            line["weak"] = weak
        if world == 1 and not args.no_cpu:
            cpu = host_cpu_info()
            threads = cpu["threads"]
            n_sample = args.cpu_sample or max(threads, min(batch.n_blocks, 64 * threads))
            sub, secs, kind, simd, hashes = cpu_reference(batch, params_kw, n_sample, threads, want_hash=last is not None)
            if last is not None:
                sub_cells = float(sum(last.block(i).inband_cells for i in range(sub.n_blocks)))
            else:
                sub_cells = float(cells) * sub.n_blocks / batch.n_blocks
            line["cpu_baseline"] = {"value": sub_cells / secs / 1e9, "unit": UNIT, "cores": threads, "physical_cores": cpu["physical_cores"],
                                    "cpu_model": cpu["model"], "kind": kind, "blocks_per_s": sub.n_blocks / secs,
                                    "sample": f"first {sub.n_blocks} blocks of the batch, abPOA v1.5.4 {simd} via oracle/_ref, OpenMP dynamic over blocks, {secs:.1f} s"}
            if hashes is not None:
                # every sampled block's graph: the unmodified abPOA's hash against the hash of this run's GPU result
                bad = [i for i in range(sub.n_blocks) if int(hashes[i]) != last.block_hash(i)]
                line["parity_sample"] = {"blocks": sub.n_blocks, "mismatches": len(bad), "against": f"unmodified abPOA v1.5.4 ({simd})",
                                         "what": "FNV-1a of node count, bases, out-edge ids and weights per block"}
                if bad:
                    emit(line)
                    raise SystemExit(f"bench.py: GPU result differs from the reference on sampled blocks {bad[:8]}")
            if kind == "reference":
                n2 = max(threads, sub.n_blocks // 2)
                s2, secs2, _, simd2, _ = cpu_reference(batch, params_kw, n2, threads, isa="avx2")
                if secs2:
                    c2 = float(sum(last.block(i).inband_cells for i in range(s2.n_blocks))) if last is not None else float(cells) * s2.n_blocks / batch.n_blocks
                    line["cpu_baseline"]["avx2"] = {"value": c2 / secs2 / 1e9, "unit": UNIT, "blocks_per_s": s2.n_blocks / secs2, "simd": simd2,
                                                    "blocks": s2.n_blocks}
        if world == 1 and not args.no_e2e and args.workload == "10000x32x2kb":
            line["next_rows"] = {"adaptive_identity_estimate": identity_estimate_leg(batch, local_rank)}
        emit(line)
        if verify is not None and verify["differing_from_single_gpu_run"]:
            raise SystemExit("bench.py: the gathered multi-GPU result differs from the single-GPU run")
    if last is not None:
        last.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
