set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r06_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r06_pytest_gpu.log
tail -5 gpurun_out/r06_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r06_bench.json 2> gpurun_out/r06_bench.err
cat gpurun_out/r06_bench.json
for mb in 16 20; do POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_mb$mb.so python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r06_variants.jsonl 2>> gpurun_out/r06_variants.err; done
POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_mb16.so python bench.py --ctas-per-sm 16 --steps 1 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r06_variants.jsonl 2>> gpurun_out/r06_variants.err
POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_mb20.so python bench.py --ctas-per-sm 20 --steps 1 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r06_variants.jsonl 2>> gpurun_out/r06_variants.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/r06_full python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r06_ncu_full.log 2>&1
ls -la gpurun_out
