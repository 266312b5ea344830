# Run bench.py (device-resident leg only) once per alternative build of libpoa_b200.so found in smoothxg_b200/lib/variants/.
# File name convention: libpoa_<tag>_c<ctas-per-sm>.so.  usage: bash scripts/gpu_variants.sh <out-prefix> [extra bench args]
pre=$1; shift
mkdir -p gpurun_out
for so in smoothxg_b200/lib/variants/*.so; do
  tag=$(basename $so .so); c=${tag##*_c}
  POA_B200_LIB=$so python bench.py --ctas-per-sm $c --steps 2 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/${pre}_${tag}.json 2> gpurun_out/${pre}_${tag}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${pre}_${tag}.json"))
    pc = d["engine"]["phase_cycles"]; tot = pc["total"]
    print("VARIANT ${tag}", round(d["value"], 1), "Gcells/s", round(d["ms_per_step"], 1), "ms; fill share", round(pc["fill"] / tot, 3), "ctas", d["engine"]["n_ctas"], "retried", d["engine"]["retried_blocks"])
except Exception as e:
    print("VARIANT ${tag} failed:", e)
PY
done
