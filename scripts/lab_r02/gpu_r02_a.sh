# round 2, call A: parity of the two-chunk fill on the GPU, then the ILP / register-budget variants of the packed fill
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
bash scripts/gpu_variants.sh r02a
