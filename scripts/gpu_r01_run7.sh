set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r07_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r07_pytest_gpu.log
tail -5 gpurun_out/r07_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r07_bench.json 2> gpurun_out/r07_bench.err
cat gpurun_out/r07_bench.json
POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_mb16.so python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r07_variants.jsonl 2>> gpurun_out/r07_variants.err
python bench.py --ctas-per-sm 12 --steps 1 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r07_variants.jsonl 2>> gpurun_out/r07_variants.err
python bench.py --workload 1000x16x1kb --warps 1 --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r07_variants.jsonl 2>> gpurun_out/r07_variants.err
