set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/r03_p16_full python bench.py --blocks 1184 --warps 1 --ctas-per-sm 8 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r03_ncu_full.log 2>&1
ls -la gpurun_out
