# round 2, call E (one GPU): full parity suite, sanitizer runs on the final code, the default bench line + reference arm
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_gpu.log
tail -5 gpurun_out/r02e_pytest_gpu.log; grep -E "batched|coalesced" gpurun_out/r02e_pytest_gpu.log
for w in 1 4 8; do
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_case.py $w > gpurun_out/r02e_racecheck_w$w.txt 2>&1; tail -3 gpurun_out/r02e_racecheck_w$w.txt
done
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 1 > gpurun_out/r02e_memcheck_w1.txt 2>&1; tail -2 gpurun_out/r02e_memcheck_w1.txt
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 4 > gpurun_out/r02e_memcheck_w4.txt 2>&1; tail -2 gpurun_out/r02e_memcheck_w4.txt
python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; cut -c1-2500 gpurun_out/r02e_bench.json
