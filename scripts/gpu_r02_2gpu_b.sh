# round 2, two GPUs of one box, final build: strong scaling of ONE 10 000-block batch (the driver's launch line)
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02b_2gpu_n2.json 2> gpurun_out/r02b_2gpu_n2.err
tail -2 gpurun_out/r02b_2gpu_n2.err | cut -c1-200
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02b_2gpu_n2.json') if l.startswith('{')][-1])
print('N2', round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['parts_each_step_rank0'], d.get('gather_verify'), d.get('weak'))"
