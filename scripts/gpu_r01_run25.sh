# r25: resident-warp sweep of the committed kernel (is throughput linear in warps per SM?) and an ncu capture of a variant
set -x
mkdir -p gpurun_out
for c in 4 8 12; do
  python bench.py --ctas-per-sm $c --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r25_c$c.json 2> gpurun_out/r25_c$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r25_c$c.json")); print("ctas_per_sm $c", round(d["value"],1), d["engine"]["n_ctas"])
PY
done
v=$1
if [ -n "$v" ]; then
export POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_$v.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/r25_$v python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r25_ncu.log 2>&1
ls -la gpurun_out/r25_$v.ncu-rep
fi
