# round 2: chunk pass specialised on "profile staged" (no per-pass select between shared and global profile), band-edge flags per row; ring of 4 chunks
set -x
bash scripts/gpu_variants.sh r02y
POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_k2_c16.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02y_pytest_parity.log 2>&1; tail -3 gpurun_out/r02y_pytest_parity.log
