# Round-2 final measurements on one B200 (what the driver runs at round end, plus the evidence kept under profiles/):
# parity suite, smoke, the default bench line + the reference arm, the ncu launch list of the same command, one ncu --set full
# capture of the POA kernel (one wave of the benchmark shape), and the lines of the other BASELINE configs with their CPU arms.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
tail -3 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log; tail -3 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
cut -c1-400 gpurun_out/final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
cut -c1-300 gpurun_out/final_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poa_b200_block -c 1 -o gpurun_out/final_full python bench.py --blocks 2368 --warps 1 --ctas-per-sm 16 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/final_ncu_full.log 2>&1
python bench.py --workload 1000x16x1kb --steps 3 --warmup 3 > gpurun_out/final_config1.json 2> gpurun_out/final_config1.err
python bench.py --workload 100x256x8kb --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/final_config3.json 2> gpurun_out/final_config3.err
python -c "
import json
for w in ('config1','config3'):
    d=json.load(open('gpurun_out/final_%s.json'%w)); print(w, round(d['value'],1), round(d['e2e']['value'],1), d['cpu_baseline']['value'], d.get('parity_sample'))"
ls -la gpurun_out | tail -12
