set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err
cat gpurun_out/r19_bench.json
POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_u2.so python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r19_u2.json 2>> gpurun_out/r19_bench.err
cat gpurun_out/r19_u2.json
compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r19_racecheck.log 2>&1; tail -5 gpurun_out/r19_racecheck.log
python -m pytest tests -m gpu -x -q > gpurun_out/r19_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r19_pytest_gpu.log
tail -3 gpurun_out/r19_pytest_gpu.log
