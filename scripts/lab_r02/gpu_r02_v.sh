# round 2: row-loop instruction diet, second batch (profile wait hoisted out of the chunk pass, 32-bit slab bookkeeping, one-instruction
# metadata gather over 16-byte windows, pinned profile pointer)
set -x
bash scripts/gpu_variants.sh r02v
POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_g8_c16.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02v_pytest_parity.log 2>&1; tail -3 gpurun_out/r02v_pytest_parity.log
