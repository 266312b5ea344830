# round 2: packed fill instantiated for smoothxg's default scoring (constants as immediates) against the generic build
set -x
bash scripts/gpu_variants.sh r02p
POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_preset_c16.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02p_pytest_parity.log 2>&1; tail -3 gpurun_out/r02p_pytest_parity.log
