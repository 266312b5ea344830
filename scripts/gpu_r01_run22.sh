# r22: skip-fast traceback + warp-sweep remain: parity then bench (device-only, no cpu leg)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r22_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r22_pytest_gpu.log
tail -3 gpurun_out/r22_pytest_gpu.log
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err
cat gpurun_out/r22_bench.json
