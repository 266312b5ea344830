# r24: parity + device-only bench of a library variant (POA_B200_LIB override)
set -x
mkdir -p gpurun_out
v=$1
export POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_$v.so
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r24_${v}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r24_${v}_pytest.log
tail -3 gpurun_out/r24_${v}_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r24_$v.json 2> gpurun_out/r24_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/r24_$v.json")); print("$v", round(d["value"],1), d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["engine"]["phase_cycles"])
PY
