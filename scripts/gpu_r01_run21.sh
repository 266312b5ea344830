set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r21_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r21_pytest_gpu.log
tail -4 gpurun_out/r21_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r21_bench.json 2> gpurun_out/r21_bench.err; cat gpurun_out/r21_bench.json
for w in 1 2 4 8; do python bench.py --workload 1000x16x1kb --blocks 148 --warps $w --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r21_small.jsonl 2>> gpurun_out/r21_bench.err; done
cat gpurun_out/r21_small.jsonl | cut -c1-400
timeout 900 python bench.py --workload 100x256x8kb --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r21_deep.json 2> gpurun_out/r21_deep.err; echo "rc=$?" >> gpurun_out/r21_deep.err
cat gpurun_out/r21_deep.json | cut -c1-2000; tail -2 gpurun_out/r21_deep.err
