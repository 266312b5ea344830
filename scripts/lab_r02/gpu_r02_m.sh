# round 2: multi-warp packed fill with the single-warp fill's gather (one zero-filling copy over 16-byte windows) and a
# warp-parallel combine of the per-warp arg-max candidates; deep blocks (configs[3]) and a small batch (300 x 16 x 1 kb, 4 warps)
set -x
bash scripts/gpu_variants.sh r02m3 --workload 100x256x8kb --steps 1 --warmup 1
bash scripts/gpu_variants.sh r02m1 --workload 1000x16x1kb --blocks 300 --warps 4 --steps 3 --warmup 2
POA_B200_LIB=smoothxg_b200/lib/variants/libpoa_mw2_c1.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02m_pytest_parity.log 2>&1; tail -3 gpurun_out/r02m_pytest_parity.log
