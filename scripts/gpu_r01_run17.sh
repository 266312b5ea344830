set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 1 --no-cpu > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
cat gpurun_out/r17_bench.json
python -m pytest tests -m gpu -x -q > gpurun_out/r17_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r17_pytest_gpu.log
tail -3 gpurun_out/r17_pytest_gpu.log
