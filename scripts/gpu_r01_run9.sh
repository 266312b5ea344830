set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r09_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r09_pytest_gpu.log
tail -5 gpurun_out/r09_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r09_bench.json 2> gpurun_out/r09_bench.err
cat gpurun_out/r09_bench.json
POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_nopeel.so python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r09_variants.jsonl 2>> gpurun_out/r09_variants.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/r09_full python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r09_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
