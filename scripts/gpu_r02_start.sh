# First GPU call of the next round (nothing here has been measured yet): what v12 changed for the deep-block workload,
# small-batch warps-per-block choice, and the headline line of the committed build.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r02_start.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest_gpu.log
tail -3 gpurun_out/r02s_pytest_gpu.log
# configs[3]: 8.1 Gcells/s at v11 (five-plane generic rows, 81 of 100 blocks resident); v12 stores three planes
python bench.py --workload 100x256x8kb --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02s_config3.json 2> gpurun_out/r02s_config3.err
python -c "import json; d=json.load(open('gpurun_out/r02s_config3.json')); print('CONFIG3', round(d['value'],1), d['engine'])"
for w in 4 8; do
  python bench.py --workload 100x256x8kb --warps $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02s_config3_w$w.json 2> gpurun_out/r02s_config3_w$w.err
  python -c "import json; d=json.load(open('gpurun_out/r02s_config3_w$w.json')); print('CONFIG3 warps $w', round(d['value'],1), d['engine']['n_ctas'], d['engine']['workspace_gb'])"
done
# configs[1]: 1 000 blocks fill 42 % of the resident block slots at one warp per block (145 Gcells/s at v11)
for w in 1 2 4; do
  python bench.py --workload 1000x16x1kb --warps $w --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02s_config1_w$w.json 2> gpurun_out/r02s_config1_w$w.err
  python -c "import json; d=json.load(open('gpurun_out/r02s_config1_w$w.json')); print('CONFIG1 warps $w', round(d['value'],1), round(d['blocks_per_s'],1))"
done
# SURVEY 8(d) variants of the headline batch: the -a preset regimes (0.1 % and 5 % divergence) and long indels
for w in 10000x32x2kb_d0.1 10000x32x2kb_d5 10000x32x2kb_indel; do
  python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02s_$w.json 2> gpurun_out/r02s_$w.err
  python -c "import json; d=json.load(open('gpurun_out/r02s_$w.json')); print('VARIANT $w', round(d['value'],1), round(d['blocks_per_s'],1), d['engine']['retried_blocks'], round(d['p_bar'],3))"
done
python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
cat gpurun_out/r02s_bench.json
