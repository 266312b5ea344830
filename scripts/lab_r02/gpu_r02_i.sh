# round 2, call I (one GPU): four-predecessor metadata prefetch in both packed fills, occupancy-derived CTA counts
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log; tail -4 gpurun_out/r02i_pytest.log; grep -E "^blocks " gpurun_out/r02i_pytest.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02i_config2.json 2> gpurun_out/r02i_config2.err
python -c "import json; d=json.load(open('gpurun_out/r02i_config2.json')); print('CONFIG2', round(d['value'],1), round(d['ms_per_step'],1), d['engine']['n_ctas'], d['engine']['warps_per_block'])"
python bench.py --workload 100x256x8kb --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02i_config3.json 2> gpurun_out/r02i_config3.err
python -c "import json; d=json.load(open('gpurun_out/r02i_config3.json')); print('CONFIG3', round(d['value'],1), round(d['ms_per_step']), d['engine']['n_ctas'], d['engine']['warps_per_block'], round(d['engine']['workspace_gb'],1), d['engine']['retried_blocks'])"
python bench.py --workload 1000x16x1kb --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02i_config1.json 2> gpurun_out/r02i_config1.err
python -c "import json; d=json.load(open('gpurun_out/r02i_config1.json')); print('CONFIG1', round(d['value'],1), round(d['ms_per_step'],1), round(d['roofline']['kernel_ms_per_launch'],1), d['engine']['n_ctas'], d['engine']['warps_per_block'])"
for nb in 300 140; do
  python bench.py --blocks $nb --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02i_b$nb.json 2> gpurun_out/r02i_b$nb.err
  python -c "import json; d=json.load(open('gpurun_out/r02i_b$nb.json')); print('BLOCKS $nb', round(d['value'],1), round(d['ms_per_step'],1), d['engine']['n_ctas'], d['engine']['warps_per_block'])"
done
