# r27: identity-estimate kernels (parity + timing), "is the fill bound by DRAM writes?" (a sixth never-read plane per row),
# write-only vs copy bandwidth of the box, and where the end-to-end second goes.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_mash.py -m gpu -x -q > gpurun_out/r27_pytest_mash.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r27_pytest_mash.log
tail -15 gpurun_out/r27_pytest_mash.log
python scripts/bench_mash.py > gpurun_out/r27_bench_mash.json 2> gpurun_out/r27_bench_mash.err; cat gpurun_out/r27_bench_mash.json; tail -3 gpurun_out/r27_bench_mash.err
python - <<'PY' > gpurun_out/r27_bw.txt 2>&1
import torch
x = torch.empty(1 << 32, dtype=torch.uint8, device="cuda"); y = torch.empty_like(x)
def t(f, n=10):
    f(); torch.cuda.synchronize(); best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); f(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_()); print("memset GB/s", x.numel() / ms / 1e6)
ms = t(lambda: y.copy_(x)); print("copy GB/s (r+w)", 2 * x.numel() / ms / 1e6)
ms = t(lambda: x.sum(dtype=torch.int64)) if False else 0
z = x.view(torch.int64)
ms = t(lambda: torch.sum(z)); print("read GB/s", x.numel() / ms / 1e6)
PY
cat gpurun_out/r27_bw.txt
L=$PWD/smoothxg_b200/lib
for v in libpoa_b200.so libpoa_b200_vH.so; do
  POA_B200_LIB=$L/$v python bench.py --ctas-per-sm 12 --slab-rows-factor 2.2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r27_$v.json 2> gpurun_out/r27_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r27_$v.json")); print("VARIANT $v", round(d["value"], 1), d["engine"]["n_ctas"], d["engine"]["retried_blocks"], d["engine"]["workspace_gb"])
PY
done
python scripts/e2e_breakdown.py > gpurun_out/r27_e2e_breakdown.txt 2>&1; cat gpurun_out/r27_e2e_breakdown.txt | tail -8
