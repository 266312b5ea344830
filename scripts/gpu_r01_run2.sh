set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
cat gpurun_out/r02_bench.json
for v in "1 8" "1 24"; do set -- $v; python bench.py --warps $1 --ctas-per-sm $2 --steps 1 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r02_variants.jsonl 2>> gpurun_out/r02_variants.err; done
python bench.py --workload 1000x16x1kb --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r02_variants.jsonl 2>> gpurun_out/r02_variants.err
python bench.py --workload 1000x16x1kb --warps 1 --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r02_variants.jsonl 2>> gpurun_out/r02_variants.err
