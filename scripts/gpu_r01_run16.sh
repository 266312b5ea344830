set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err
cat gpurun_out/r16_bench.json
python -m pytest tests -m gpu -x -q > gpurun_out/r16_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r16_pytest_gpu.log
tail -3 gpurun_out/r16_pytest_gpu.log
python bench.py --workload 1000x16x1kb --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r16_cfg1_auto.json 2>> gpurun_out/r16_bench.err
python bench.py --workload 1000x16x1kb --warps 1 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r16_cfg1_w1.json 2>> gpurun_out/r16_bench.err
