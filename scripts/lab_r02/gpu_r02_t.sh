# round 2: profile chunks by one TMA bulk copy per row (mbarrier), branch-free metadata gather, one-chunk arg-max epilogue
set -x
bash scripts/gpu_variants.sh r02t
POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_qtmagflateq_c16.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02t_pytest_parity.log 2>&1; tail -3 gpurun_out/r02t_pytest_parity.log
