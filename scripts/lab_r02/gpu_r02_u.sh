# round 2: row-loop instruction diet, cumulative variants (branch-free gather; bounded profile staging + one-chunk arg-max; profile
# group committed before the gather; lane constants from shared memory + pinned -inf fill)
set -x
bash scripts/gpu_variants.sh r02u
POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_g4_c16.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02u_pytest_parity.log 2>&1; tail -3 gpurun_out/r02u_pytest_parity.log
