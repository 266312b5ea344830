# round 2, last refresh of the headline line and its launch list on the final build (default-scoring preset included)
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err
cut -c1-300 gpurun_out/final3_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/final3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final3_launches_bench.log 2>&1
python bench.py --workload 1000x16x1kb --steps 3 --warmup 3 --no-cpu > gpurun_out/final3_config1.json 2> gpurun_out/final3_config1.err
python -c "
import json
d=json.load(open('gpurun_out/final3_bench.json')); print('FINAL', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), d['cpu_baseline']['value'], d['parity_sample'])
d=json.load(open('gpurun_out/final3_config1.json')); print('C1', round(d['value'],1), round(d['e2e']['value'],1))"
