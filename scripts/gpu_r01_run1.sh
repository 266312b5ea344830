set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r01_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/r01_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r01_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01_pytest_gpu.log
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
for v in "2 8" "1 32" "4 4" "1 8"; do set -- $v; python bench.py --warps $1 --ctas-per-sm $2 --steps 1 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r01_variants.jsonl 2>> gpurun_out/r01_variants.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r01_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/r01_poa_full python bench.py --blocks 1184 --ctas-per-sm 8 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r01_ncu_full.log 2>&1
ls -la gpurun_out
