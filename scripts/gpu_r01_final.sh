# Round-1 final measurements on one B200: parity tests, smoke, the bench line, the ncu launch list of the same
# command, and one ncu --set full capture of the POA kernel (one wave of blocks of the benchmark shape).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
tail -3 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
cat gpurun_out/final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
cat gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/final_full python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
