# round 2, call G (one GPU): deep blocks on the packed 16-bit fill (p16_safe_for_long_graph): parity of the full-size block, then configs[3]
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_single or int32 or deep or golden" > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log; tail -4 gpurun_out/r02g_pytest.log
for w in 8 4 2; do
  python bench.py --workload 100x256x8kb --warps $w --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02g_config3_w$w.json 2> gpurun_out/r02g_config3_w$w.err
  python -c "import json; d=json.load(open('gpurun_out/r02g_config3_w$w.json')); print('CONFIG3 warps $w', round(d['value'],1), round(d['ms_per_step']), d['engine']['n_ctas'], round(d['engine']['workspace_gb'],1), d['engine']['retried_blocks'], d['engine']['phase_cycles']['spare'])"
done
