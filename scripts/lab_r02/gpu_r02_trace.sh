# host-side wall time of the stages of poa_b200_run_batch on the small configuration (where it shows next to a 56 ms kernel)
set -x
POA_B200_TRACE=1 python bench.py --workload 1000x16x1kb --steps 3 --warmup 3 --no-cpu > gpurun_out/trace_config1.json 2> gpurun_out/trace_config1.err
grep "poa_b200" gpurun_out/trace_config1.err | tail -8
python -c "
import json; d=json.load(open('gpurun_out/trace_config1.json')); print('E2E', d['e2e']['ms_each_step_rank0'], d['e2e']['parts_each_step_rank0'], d['value'], d['e2e']['value'])"
