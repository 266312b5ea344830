set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r20_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r20_pytest_gpu.log
tail -4 gpurun_out/r20_pytest_gpu.log
timeout 900 python bench.py --workload 100x256x8kb --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r20_deep.json 2> gpurun_out/r20_deep.err; echo "rc=$?" >> gpurun_out/r20_deep.err
cat gpurun_out/r20_deep.json; tail -3 gpurun_out/r20_deep.err
