# round 2, last call: whole GPU suite on the final build, configs[3] and configs[1] lines of the final build
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/final5_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final5_pytest_gpu.log
tail -3 gpurun_out/final5_pytest_gpu.log
python bench.py --workload 100x256x8kb --steps 1 --warmup 1 --no-cpu > gpurun_out/final5_config3.json 2> gpurun_out/final5_config3.err
python bench.py --workload 1000x16x1kb --steps 3 --warmup 3 --no-cpu > gpurun_out/final5_config1.json 2> gpurun_out/final5_config1.err
python -c "
import json
for w in ('config1','config3'):
    d=json.load(open('gpurun_out/final5_%s.json'%w)); print(w, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), d['engine']['warps_per_block'])"
