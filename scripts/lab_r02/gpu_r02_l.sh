# round 2, call L (one GPU): multi-warp fill with all loads of a chunk issued up front: parity of the multi-warp paths, deep blocks, small batches
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log; tail -3 gpurun_out/r02l_pytest.log
python bench.py --workload 100x256x8kb --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02l_config3.json 2> gpurun_out/r02l_config3.err
python -c "import json; d=json.load(open('gpurun_out/r02l_config3.json')); print('CONFIG3', round(d['value'],1), round(d['ms_per_step']), d['engine']['warps_per_block'])"
for nb in 300 140; do
  python bench.py --blocks $nb --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02l_b$nb.json 2> gpurun_out/r02l_b$nb.err
  python -c "import json; d=json.load(open('gpurun_out/r02l_b$nb.json')); print('BLOCKS $nb', round(d['value'],1), round(d['ms_per_step'],1), d['engine']['n_ctas'], d['engine']['warps_per_block'])"
done
