# round 2, two GPUs of one box: strong scaling of ONE 10 000-block batch (LPT shards + NCCL gather, bit-identity check against the
# single-GPU run inside bench.py), next to the N = 1 line of the same box; then the small sharded parity check
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/r02_2gpu_n1.json 2> gpurun_out/r02_2gpu_n1.err
python -c "import json; d=json.load(open('gpurun_out/r02_2gpu_n1.json')); print('N1', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r02_2gpu_n2.json 2> gpurun_out/r02_2gpu_n2.err
tail -3 gpurun_out/r02_2gpu_n2.err
python -c "import json; d=json.load(open('gpurun_out/r02_2gpu_n2.json')); print('N2', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['e2e']['parts_each_step_rank0'], d.get('gather_verify'), d.get('weak'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 scripts/sharded_2gpu_check.py > gpurun_out/r02_2gpu_sharded_check.log 2>&1; tail -2 gpurun_out/r02_2gpu_sharded_check.log
