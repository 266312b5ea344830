# r32: GPU parity of the build whose generic fill stores H / E planes only (F recomputed in the traceback for all gap modes and int32)
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r32_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r32_pytest.log
tail -5 gpurun_out/r32_pytest.log
