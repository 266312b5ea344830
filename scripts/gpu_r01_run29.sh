# r29: two small variants of the v11 kernel (profile staged one row ahead; evict-first stores of the row planes), parity of
# the best one, and the two other BASELINE workloads (configs[1], configs[3]) for the record.
set -x
mkdir -p gpurun_out
L=$PWD/smoothxg_b200/lib
for v in libpoa_b200.so libpoa_b200_vQ.so libpoa_b200_vS.so libpoa_b200_vQS.so; do
  POA_B200_LIB=$L/$v python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r29_$v.json 2> gpurun_out/r29_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r29_$v.json")); print("VARIANT $v", round(d["value"], 1), d["engine"]["n_ctas"], d["engine"]["retried_blocks"])
except Exception as e:
    print("VARIANT $v failed", e)
PY
done
python - > gpurun_out/r29_best.txt <<'PY'
import json
best = None
for v in ("libpoa_b200_vQ.so", "libpoa_b200_vS.so", "libpoa_b200_vQS.so"):
    try:
        x = json.load(open(f"gpurun_out/r29_{v}.json"))["value"]
    except Exception:
        continue
    if best is None or x > best[0]:
        best = (x, v)
print(best[1])
PY
read BLIB < gpurun_out/r29_best.txt
echo "BEST $BLIB"
POA_B200_LIB=$L/$BLIB python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r29_pytest_best.log 2>&1; echo "pytest rc=$? lib=$BLIB" >> gpurun_out/r29_pytest_best.log
tail -3 gpurun_out/r29_pytest_best.log
for w in 1000x16x1kb 100x256x8kb; do
  python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r29_w_$w.json 2> gpurun_out/r29_w_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r29_w_$w.json")); print("WORKLOAD $w", round(d["value"], 1), "Gcells/s", round(d["blocks_per_s"], 1), "blocks/s", d["engine"])
except Exception as e:
    print("WORKLOAD $w failed", e)
PY
done
