# round 2, after the row-loop diet: racecheck + memcheck of the single-warp kernel (the code that changed), GPU fuzz with fresh seeds
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_case.py 1 > gpurun_out/r02s_racecheck_w1.txt 2>&1; tail -3 gpurun_out/r02s_racecheck_w1.txt
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 1 > gpurun_out/r02s_memcheck_w1.txt 2>&1; tail -2 gpurun_out/r02s_memcheck_w1.txt
timeout 500 python scripts/fuzz_gpu.py 40 20261017 > gpurun_out/r02s_fuzz_gpu.log 2>&1; tail -3 gpurun_out/r02s_fuzz_gpu.log
