# r23: A/B of library variants named on the command line (POA_B200_LIB override), device-only bench
set -x
mkdir -p gpurun_out
for v in "$@"; do
  POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_$v.so python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r23_$v.json 2> gpurun_out/r23_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r23_$v.json")); print("$v", round(d["value"],1), d["engine"]["phase_cycles"])
PY
done
