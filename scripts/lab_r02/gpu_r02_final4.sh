# round 2: the whole GPU suite and smoke() once more on the final build
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/final4_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final4_pytest_gpu.log
tail -3 gpurun_out/final4_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final4_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final4_smoke.log; tail -3 gpurun_out/final4_smoke.log
