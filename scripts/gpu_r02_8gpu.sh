# round 2, eight GPUs of one box: the strong-scaling line exactly as the driver launches it (one 10 000-block batch, LPT shards, NCCL gather)
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_8gpu_n8.json 2> gpurun_out/r02_8gpu_n8.err
tail -3 gpurun_out/r02_8gpu_n8.err | cut -c1-300
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_8gpu_n8.json') if l.startswith('{')][-1])
print('N8', round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],1), d['e2e']['parts_each_step_rank0'], d.get('gather_verify'), d.get('weak'))"
