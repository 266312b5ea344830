set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/g2_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_2gpu_check.py > gpurun_out/g2_sharded.log 2>&1; echo "rc=$?" >> gpurun_out/g2_sharded.log
tail -3 gpurun_out/g2_sharded.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; echo "rc=$?" >> gpurun_out/g2_bench.err
cat gpurun_out/g2_bench.json; tail -3 gpurun_out/g2_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/g2_bench_ref.json 2>> gpurun_out/g2_bench.err
cat gpurun_out/g2_bench_ref.json
