# round 2: first-predecessor window in shared memory for the traceback; profile chunks requested one row ahead; row maximum before the stores
set -x
bash scripts/gpu_variants.sh r02x
POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_h3_c16.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02x_pytest_parity.log 2>&1; tail -3 gpurun_out/r02x_pytest_parity.log
python - <<'PY'
import json
for t in ("g8","h1","h2","h3"):
    d=json.load(open(f"gpurun_out/r02x_libpoa_{t}_c16.json")); pc=d["engine"]["phase_cycles"]; tot=pc["total"]
    print("PHASES",t,{k:round(v/tot,4) for k,v in pc.items()})
PY
