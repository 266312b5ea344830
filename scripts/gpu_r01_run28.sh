# r28: H/E1/E2-only rows (F planes recomputed on the traceback's insertion steps): parity, smoke, the bench line with e2e and
# CPU baseline, the reference arm, the ncu launch list of the bench command and a sectioned ncu capture of one wave.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r28_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r28_pytest_gpu.log
tail -4 gpurun_out/r28_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r28_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r28_smoke.log
tail -2 gpurun_out/r28_smoke.log
python bench.py > gpurun_out/r28_bench.json 2> gpurun_out/r28_bench.err
cat gpurun_out/r28_bench.json; tail -3 gpurun_out/r28_bench.err
python scripts/e2e_breakdown.py > gpurun_out/r28_e2e_breakdown.txt 2>&1; tail -6 gpurun_out/r28_e2e_breakdown.txt
python bench.py --ctas-per-sm 12 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r28_c12.json 2> gpurun_out/r28_c12.err
python -c "import json; d=json.load(open('gpurun_out/r28_c12.json')); print('CTAS12', round(d['value'],1), d['engine']['workspace_gb'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r28_bench_reference.json 2>> gpurun_out/r28_bench.err
cat gpurun_out/r28_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/r28_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r28_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/r28_full python bench.py --blocks 2368 --warps 1 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r28_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
