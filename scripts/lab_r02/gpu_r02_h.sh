# round 2, call H (one GPU): the reworked multi-warp packed fill (metadata prefetch, one barrier per row + one per round)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log; tail -4 gpurun_out/r02h_pytest.log
for w in 8 4; do
  python bench.py --workload 100x256x8kb --warps $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02h_config3_w$w.json 2> gpurun_out/r02h_config3_w$w.err
  python -c "import json; d=json.load(open('gpurun_out/r02h_config3_w$w.json')); print('CONFIG3 warps $w', round(d['value'],1), round(d['ms_per_step']), d['engine']['n_ctas'], round(d['engine']['workspace_gb'],1), d['engine']['retried_blocks'])"
done
for w in 1 2 4; do
  python bench.py --workload 1000x16x1kb --warps $w --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02h_config1_w$w.json 2> gpurun_out/r02h_config1_w$w.err
  python -c "import json; d=json.load(open('gpurun_out/r02h_config1_w$w.json')); print('CONFIG1 warps $w', round(d['value'],1), round(d['ms_per_step'],1), round(d['roofline']['kernel_ms_per_launch'],1))"
done
for nb in 1250 2500; do for w in 1 2; do
  python bench.py --blocks $nb --warps $w --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02h_b${nb}_w$w.json 2> gpurun_out/r02h_b${nb}_w$w.err
  python -c "import json; d=json.load(open('gpurun_out/r02h_b${nb}_w$w.json')); print('BLOCKS $nb warps $w', round(d['value'],1), round(d['ms_per_step'],1), d['engine']['n_ctas'])"
done; done
